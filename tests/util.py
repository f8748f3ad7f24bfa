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


def oracle_eval_chunked(p, spec, x, b, keys, K, rows=128, dtype=torch.float64, want_imp=True):
    """oracle eval_fn (+ log p(x)) over B rows in chunks of `rows`, eps drawn from the three JAX keys exactly as the
    device draws them (normal(key, [K, B, d]) restricted to the chunk's rows): the [K * B, 256] activations of the
    decoder never exist at once."""
    from oracle import prng
    B = x.shape[0]
    k_imp, k_z, k_zxo = keys
    imp, ll, lpx = [], [], []
    pd = M.cast_params(p, dtype)
    for r0 in range(0, B, rows):
        nb = min(rows, B - r0)
        e = []
        for k in ((k_imp, k_z, k_zxo) if want_imp else (k_z, k_zxo)):
            full = prng.normal_rows(k, K, B, spec.d, r0, nb)
            e.append(torch.tensor(full, dtype=dtype))
        xs, bs = x[r0:r0 + nb].to(dtype), b[r0:r0 + nb].to(dtype)
        with torch.no_grad():
            if want_imp:
                imp.append(M.impute(pd, spec, xs, bs, e[0]).mean(0))
            a, c = M.is_log_prob(pd, spec, xs, bs, e[-2], e[-1])
        lpx.append(a)
        ll.append(c)
    return (torch.cat(imp) if want_imp else None), torch.cat(ll), torch.cat(lpx)
