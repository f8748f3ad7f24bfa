"""CPU restatement (PyTorch float64) of the Posterior Matching term of PosteriorMatchingVADE
(posterior_matching/models/vade.py:246-265; train_pm_vade.py:38-41,57-58) for configs/pm_vade_mnist.py: ConvEncoder +
DiagonalGaussian posterior (distributions.py:58-84), ConvEncoder partial encoder + AutoregressiveGMM.  TEST INFRASTRUCTURE
ONLY.  Haiku leaf names are [R] (the DiagonalGaussian module is unnamed in vade.py:61-63 -> `diagonal_gaussian`)."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import conv as OC
from . import dists_mnist as DM

LATENT = 10
SPEC = DM.ArgmmSpec(d=LATENT, n_comp=10, R=2, H=256, C=128)


def _name(prefix, base, i):
    return f"{prefix}/{base}" if i == 0 else f"{prefix}/{base}_{i}"


def leaf_shapes():
    out = []
    cin = 1
    for i, (f, k, _) in enumerate(OC.MNIST_ENCODER):
        out.append((_name("encoder_net", "conv2_d", i), (k, k, cin, f), f)); cin = f
    out.append(("diagonal_gaussian/linear", (128, 2 * LATENT), 2 * LATENT))
    cin = 2
    for i, (f, k, _) in enumerate(OC.MNIST_ENCODER):
        out.append((_name("partial_encoder_net", "conv2_d", i), (k, k, cin, f), f)); cin = f
    for n, fi, fo in DM.argmm_leaf_shapes(SPEC):
        out.append((n, (fi, fo), fo))
    return out


def init_params(seed: int = 11, dtype=torch.float64):
    rng = np.random.default_rng(seed)
    p = {}
    for name, wshape, nb in leaf_shapes():
        fan_in = int(np.prod(wshape[:-1]))
        w = np.clip(rng.standard_normal(wshape), -2, 2) / math.sqrt(fan_in)
        p[name] = {"w": torch.tensor(w.astype(np.float32), dtype=dtype),
                   "b": torch.tensor((0.05 * rng.standard_normal(nb)).astype(np.float32), dtype=dtype)}
    return p


def _convs(p, prefix, n):
    return [(p[_name(prefix, "conv2_d", i)]["w"], p[_name(prefix, "conv2_d", i)]["b"]) for i in range(n)]


def posterior_matching_ll(p, x, b, eps):
    """vade.py:246-265: z = loc + (softplus(raw) + 1e-5) eps, log q(stop_gradient(z) | x_o)."""
    B = x.shape[0]
    h = OC.conv_encoder(_convs(p, "encoder_net", 5), x).reshape(B, -1)
    par = h @ p["diagonal_gaussian/linear"]["w"] + p["diagonal_gaussian/linear"]["b"]
    z = par[:, :LATENT] + (torch.nn.functional.softplus(par[:, LATENT:]) + 1e-5) * eps
    ctx = OC.conv_encoder(_convs(p, "partial_encoder_net", 5), torch.cat([x * b, b], -1)).reshape(B, -1)
    return DM.argmm_log_prob(p, SPEC, z.detach(), ctx), z


def loss_and_grads(p, x, b, eps):
    """train_pm_vade.py:38-41 with the trainable predicate of :57-58: gradients of -mean(ll) for `partial_*` leaves."""
    q = {n: {k: t.detach().clone().requires_grad_(n.startswith("partial_")) for k, t in leaf.items()} for n, leaf in p.items()}
    ll, _ = posterior_matching_ll(q, x, b, eps)
    loss = -ll.mean()
    loss.backward()
    return loss.detach(), ll.detach(), {n: {k: t.grad for k, t in leaf.items()} for n, leaf in q.items() if n.startswith("partial_")}
