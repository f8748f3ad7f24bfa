"""CPU restatement (PyTorch float64) of the PM-VAE of configs/pm_vae_mnist.py: ConvEncoder / ConvDecoder
(networks.py:9-72), TriLGaussian posterior (distributions.py:87-113), Bernoulli decoder (:20-25),
AutoregressiveGMM partial posterior (:137-223), PosteriorMatchingVAE.__call__ (vae.py:120-144) and loss_fn
(train_pm_vae.py:58-72; the config has no beta schedule -> beta = 1, no matching_ll_stop_gradients -> False).
TEST INFRASTRUCTURE ONLY.  Haiku leaf names ([R]): `encoder_net/conv2_d{,_1..}`, `posterior_dist/linear`,
`decoder_net/conv2_d_transpose{,_1..}`, `partial_encoder_net/conv2_d{,_1..}`, AR-GMM leaves as in dists_mnist.py.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from . import conv as OC
from . import dists_mnist as DM
from .model import fill_scale_tril, tril_kl_std_normal

LATENT = 32
P_TRIL = LATENT + LATENT * (LATENT + 1) // 2     # 560


def _name(prefix, base, i):
    return f"{prefix}/{base}" if i == 0 else f"{prefix}/{base}_{i}"


def leaf_shapes():
    out = []
    cin = 1
    for i, (f, k, _) in enumerate(OC.MNIST_ENCODER):
        out.append((_name("encoder_net", "conv2_d", i), (k, k, cin, f), f)); cin = f
    out.append(("posterior_dist/linear", (128, P_TRIL), P_TRIL))
    cin = LATENT
    for i, (f, k, _) in enumerate(OC.MNIST_DECODER):
        out.append((_name("decoder_net", "conv2_d_transpose", i), (k, k, f, cin), f)); cin = f
    cin = 2
    for i, (f, k, _) in enumerate(OC.MNIST_ENCODER):
        out.append((_name("partial_encoder_net", "conv2_d", i), (k, k, cin, f), f)); cin = f
    spec = DM.ArgmmSpec(d=LATENT, n_comp=10, R=2, H=256, C=128)
    for n, fi, fo in DM.argmm_leaf_shapes(spec):
        out.append((n, (fi, fo), fo))
    return out


def init_params(seed: int = 7, dtype=torch.float64, head_scale: float = 0.1):
    rng = np.random.default_rng(seed)
    p = {}
    for name, wshape, nb in leaf_shapes():
        fan_in = int(np.prod(wshape[:-1])) if "transpose" not in name else int(wshape[0] * wshape[1] * wshape[3])
        w = np.clip(rng.standard_normal(wshape), -2, 2) / math.sqrt(fan_in)
        if name == "posterior_dist/linear":
            w = w * head_scale
        p[name] = {"w": torch.tensor(w.astype(np.float32), dtype=dtype),
                   "b": torch.tensor((0.05 * rng.standard_normal(nb)).astype(np.float32), dtype=dtype)}
    return p


def _convs(p, prefix, base, n):
    return [(p[_name(prefix, base, i)]["w"], p[_name(prefix, base, i)]["b"]) for i in range(n)]


def forward(p, x, b, eps) -> Dict[str, torch.Tensor]:
    """x, b: [B,28,28,1]; eps [B,32] -> the three per-row terms of vae.py:120-144."""
    B = x.shape[0]
    h = OC.conv_encoder(_convs(p, "encoder_net", "conv2_d", 5), x).reshape(B, -1)
    par = h @ p["posterior_dist/linear"]["w"] + p["posterior_dist/linear"]["b"]
    mu, L = par[:, :LATENT], fill_scale_tril(par[:, LATENT:], LATENT)
    z = mu + (L @ eps.unsqueeze(-1)).squeeze(-1)
    logits = OC.conv_decoder(_convs(p, "decoder_net", "conv2_d_transpose", 6), z)
    rec = DM.bernoulli_log_prob(logits, x).reshape(B, -1).sum(-1)
    kl = tril_kl_std_normal(mu, L)
    ctx = OC.conv_encoder(_convs(p, "partial_encoder_net", "conv2_d", 5), torch.cat([x * b, b], -1)).reshape(B, -1)
    spec = DM.ArgmmSpec(d=LATENT, n_comp=10, R=2, H=256, C=128)
    match = DM.argmm_log_prob(p, spec, z, ctx)          # no stop_gradient in this config
    return {"reconstruction_ll": rec, "kl": kl, "matching_ll": match, "z": z}


def loss_fn(p, x, b, eps, beta: float = 1.0, coef: float = 1.0):
    out = forward(p, x, b, eps)
    loss = -(out["reconstruction_ll"] - beta * out["kl"]).mean() + coef * (-out["matching_ll"].mean())
    return loss, out


def loss_and_grads(p, x, b, eps, beta: float = 1.0, coef: float = 1.0):
    q = {n: {k: t.detach().clone().requires_grad_(True) for k, t in leaf.items()} for n, leaf in p.items()}
    loss, out = loss_fn(q, x, b, eps, beta, coef)
    loss.backward()
    return loss.detach(), {k: v.detach() for k, v in out.items()}, {n: {k: t.grad for k, t in leaf.items()} for n, leaf in q.items()}


def _posterior(p, x):
    B = x.shape[0]
    h = OC.conv_encoder(_convs(p, "encoder_net", "conv2_d", 5), x).reshape(B, -1)
    par = h @ p["posterior_dist/linear"]["w"] + p["posterior_dist/linear"]["b"]
    return par[:, :LATENT], fill_scale_tril(par[:, LATENT:], LATENT)


def _context(p, x, b):
    return OC.conv_encoder(_convs(p, "partial_encoder_net", "conv2_d", 5), torch.cat([x * b, b], -1)).reshape(x.shape[0], -1)


ARGMM_SPEC = DM.ArgmmSpec(d=LATENT, n_comp=10, R=2, H=256, C=128)


def impute(p, x_o, b, z_xo):
    """vae.py:146-169 given the AR-GMM samples z_xo [K,B,32] of q(z | x_o): [K,B,28,28,1], the Bernoulli mean
    (sigmoid of the decoder logits) where b == 0, x_o elsewhere."""
    K, B, d = z_xo.shape
    x_o = x_o * b
    logits = OC.conv_decoder(_convs(p, "decoder_net", "conv2_d_transpose", 6), z_xo.reshape(K * B, d))
    mean = torch.sigmoid(logits).reshape(K, *x_o.shape)
    return torch.where(b.unsqueeze(0) != 0, x_o.unsqueeze(0), mean)


def is_log_prob(p, x, b, eps_z, z_xo):
    """vae.py:171-226 for this config: eps_z [K,B,32] drives z ~ q(z|x) (TriL); z_xo [K,B,32] are samples of the
    AutoregressiveGMM partial posterior -> (log p(x), log p(x_u | x_o))."""
    from .model import std_normal_log_prob, tril_log_prob
    K, B, d = eps_z.shape
    mu, L = _posterior(p, x)
    ctx = _context(p, x, b)
    z = mu.unsqueeze(0) + torch.einsum("bij,kbj->kbi", L, eps_z)
    dec = _convs(p, "decoder_net", "conv2_d_transpose", 6)

    def dec_ll(zz, weight):
        logits = OC.conv_decoder(dec, zz.reshape(K * B, d)).reshape(K, *x.shape)
        ll = DM.bernoulli_log_prob(logits, x.unsqueeze(0))
        if weight is not None:
            ll = ll * weight.unsqueeze(0)
        return ll.reshape(K, B, -1).sum(-1)

    log_q_zgx = tril_log_prob(z, mu.unsqueeze(0), L.unsqueeze(0))
    log_q_zgxo = DM.argmm_log_prob(p, ARGMM_SPEC, z_xo.reshape(K * B, d), ctx.repeat(K, 1)).reshape(K, B)
    lk = math.log(K)
    log_p_x = torch.logsumexp(dec_ll(z, None) + std_normal_log_prob(z) - log_q_zgx, 0) - lk
    log_p_xo = torch.logsumexp(dec_ll(z_xo, b) + std_normal_log_prob(z_xo) - log_q_zgxo, 0) - lk
    return log_p_x, log_p_x - log_p_xo
