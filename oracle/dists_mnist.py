"""CPU restatement (PyTorch float64) of the distribution heads only the MNIST config uses.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference files restated (relative to /root/reference):
  posterior_matching/models/distributions.py:20-25    Bernoulli           -> `bernoulli_log_prob`
  posterior_matching/models/distributions.py:116-134  OneDimensionalGMM   -> `gmm_log_prob`
  posterior_matching/models/distributions.py:137-166  _AutoregressiveDistribution.log_prob -> `argmm_log_prob`
  posterior_matching/models/distributions.py:192-223  AutoregressiveGMM   -> `ArgmmSpec`, `argmm_leaf_shapes`
  configs/pm_vae_mnist.py:20-21 (partial_posterior_dist = AutoregressiveGMM, default config)

Third-party semantics: tfd.Bernoulli(logits).log_prob(x) with a FLOAT event [R: TFP casts the event and
computes -softplus(-l) x - softplus(l) (1 - x)], tfd.MixtureSameFamily(Categorical(logits), Normal) [V vs
torch.distributions in tests/test_oracle_mnist_dists.py], Haiku module naming inside hk.Sequential [R].
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch

from .model import LOG2PI, Params, _lin_name, linear, residual_mlp


@dataclass(frozen=True)
class ArgmmSpec:
    """AutoregressiveGMM(event_size=d, num_components, residual_blocks, hidden_units) applied to a
    flattened context of C features (distributions.py:192-223)."""
    d: int
    n_comp: int = 10
    R: int = 2
    H: int = 256
    C: int = 128

    @property
    def fan_in(self) -> int:
        return 2 * self.d + self.C

    @property
    def head_cols(self) -> int:
        return 3 * self.n_comp * self.d


NET = "partial_posterior_dist/residual_mlp"
HEAD = "partial_posterior_dist/one_dimensional_gmm/linear"


def argmm_leaf_shapes(spec: ArgmmSpec):
    out = [(_lin_name(NET, 0), spec.fan_in, spec.H)]
    out += [(_lin_name(NET, i), spec.H, spec.H) for i in range(1, 2 * spec.R + 1)]
    out.append((HEAD, spec.H, spec.head_cols))
    return out


def argmm_init(spec: ArgmmSpec, seed: int = 5, dtype=torch.float64) -> Params:
    rng = np.random.default_rng(seed)
    p: Params = {}
    for name, fi, fo in argmm_leaf_shapes(spec):
        w = np.clip(rng.standard_normal((fi, fo)), -2, 2) / math.sqrt(fi)
        p[name] = {"w": torch.tensor(w.astype(np.float32), dtype=dtype),
                   "b": torch.tensor((0.05 * rng.standard_normal(fo)).astype(np.float32), dtype=dtype)}
    return p


def gmm_log_prob(params, value, n_comp: int):
    """OneDimensionalGMM (distributions.py:124-134) evaluated at `value` [..., d]: params [..., d, 3K] ->
    log sum_k softmax(logits)_k N(value; mean_k, softplus(raw_k) + 1e-5), per dimension [..., d]."""
    logits = params[..., :n_comp]
    means = params[..., n_comp:-n_comp]
    scales = torch.nn.functional.softplus(params[..., -n_comp:]) + 1e-5
    lw = torch.log_softmax(logits, -1)
    v = value.unsqueeze(-1)
    comp = -0.5 * ((v - means) / scales) ** 2 - torch.log(scales) - 0.5 * LOG2PI
    return torch.logsumexp(lw + comp, -1)


def argmm_log_prob(p: Params, spec: ArgmmSpec, value, context):
    """_AutoregressiveDistribution.log_prob (distributions.py:152-166): for each step i the net sees
    [value * (arange(d) < i), (arange(d) < i), context] and contributes the log-density of dimension i."""
    B, d = value.shape
    ar = torch.arange(d, dtype=value.dtype)
    total = torch.zeros(B, dtype=value.dtype)
    for i in range(d):
        mask = (ar < i).to(value.dtype).expand(B, d)
        inp = torch.cat([value * mask, mask, context], -1)
        h = residual_mlp(p, NET, inp, spec.R, False)
        params = linear(p, HEAD, h).reshape(B, d, 3 * spec.n_comp)
        total = total + gmm_log_prob(params, value, spec.n_comp)[:, i]
    return total


def bernoulli_log_prob(logits, x):
    """tfd.Bernoulli(logits).log_prob(x) summed by the caller (vae.py:127-128); x may be any float in
    [0, 1] (the MNIST pipeline feeds binarised floats)."""
    sp = torch.nn.functional.softplus
    return -x * sp(-logits) - (1.0 - x) * sp(logits)


def argmm_sample(p: Params, spec: ArgmmSpec, context, n: int, key):
    """_AutoregressiveDistribution._sample_n (distributions.py:168-189) -> [n, B, d].  The reference hands the SAME key
    to `out.sample` at every step and vmaps over the batch with that one key, so the noise depends on
    (sample, dimension, component) only.  Noise contract shared with the device (include/pmvae.h, pmvae_argmm_sample)
    [R: TFP's MixtureSameFamily seed plumbing is recollection -> parity with the reference is distributional]:
    (k_comp, k_cat) = split(key); eps = normal(k_comp, [n, d, K]); u = uniform(k_cat, [n, d, K]);
    component = argmax_k(logits_k - log(-log(u))) (jax.random.categorical); x_i = mean_c + scale_c * eps_c."""
    from . import prng
    B, d, K = context.shape[0], spec.d, spec.n_comp
    ks = prng.split(key, 2)
    eps = torch.tensor(prng.normal(ks[0], (n, d, K)).astype(np.float64))
    u = prng.uniform(ks[1], (n, d, K)).astype(np.float32)
    u = np.where(u > 0, u, np.float32(1.17549435e-38))
    gumbel = torch.tensor(-np.log(-np.log(u.astype(np.float64))))
    x = torch.zeros(n, B, d, dtype=context.dtype)
    ar = torch.arange(d, dtype=context.dtype)
    ctx = context.unsqueeze(0).expand(n, B, context.shape[1]).reshape(n * B, -1)
    for i in range(d):
        mask = (ar < i).to(context.dtype).expand(n * B, d)
        inp = torch.cat([x.reshape(n * B, d) * mask, mask, ctx], -1)
        h = residual_mlp(p, NET, inp, spec.R, False)
        params = linear(p, HEAD, h).reshape(n, B, d, 3 * K)[:, :, i, :]
        logits, means = params[..., :K], params[..., K:2 * K]
        scales = torch.nn.functional.softplus(params[..., 2 * K:]) + 1e-5
        comp = torch.argmax(logits + gumbel[:, i, :].unsqueeze(1), -1, keepdim=True)
        e = eps[:, i, :].unsqueeze(1).expand(n, B, K)
        x[:, :, i] = (means.gather(-1, comp) + scales.gather(-1, comp) * e.gather(-1, comp)).squeeze(-1)
    return x
