"""Shared helpers for the parity tests (oracle <-> CUDA product)."""
import numpy as np
import torch

from oracle import model as M
from posterior_matching_b200.config import pm_vae_config


def spec_of(name):
    return M.spec_from_config(pm_vae_config(name).model.to_dict())


def conditioned_params(spec, seed=3, head_scale=0.1, perturb=True):
    """Haiku-default init; the two TriL head matrices are scaled so the triangular
    solves are well conditioned (at raw init forward substitution amplifies rounding
    by many orders of magnitude, in the reference too); biases and log_scale get small
    non-zero values so every term of every formula is exercised."""
    p = M.init_params(spec, seed)
    for hn in ("posterior_dist/linear", "partial_posterior_dist/linear"):
        p[hn]["w"] *= head_scale
    if perturb:
        g = torch.Generator().manual_seed(seed + 100)
        for n in p:
            if "b" in p[n]:
                p[n]["b"] = 0.05 * torch.randn(p[n]["b"].shape, generator=g, dtype=torch.float64)
        p["decoder_dist"]["log_scale"] = torch.tensor(-0.3, dtype=torch.float64)
    return p


def make_inputs(spec, B, seed=0):
    from oracle import prng
    rng = np.random.default_rng(seed)
    x = torch.tensor(rng.standard_normal((B, spec.D)).astype(np.float32), dtype=torch.float64)
    b = torch.tensor(prng.bernoulli(prng.PRNGKey(seed + 1), 0.5, (B, spec.D)).astype(np.float64))
    eps = torch.tensor(prng.normal(prng.PRNGKey(seed + 2), (B, spec.d)).astype(np.float64))
    return x, b, eps


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
