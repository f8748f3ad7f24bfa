"""CPU restatement (PyTorch float64 + the NumPy threefry stream) of `LookaheadPosterior.__call__` and
`expected_info_gains` (posterior_matching/models/lookahead.py:122-227) over the ResidualMLP / TriLGaussian /
IdentityGaussian PM-VAE of oracle/model.py, and of the loss of train_lookahead_posterior.py:47-52.
TEST INFRASTRUCTURE ONLY.  Parity unpinned against JAX itself (no JAX in this image): the Haiku key order and
`jax.random.choice(replace=False)` are restated from the sources named in prng.permutation / below [R].

Key order of one call (`hk.next_rng_key()`; every ResidualMLP block draws a dropout key even at rate 0,
networks.py:124):  R_part | z | R_dec | choice | split(K) | R_part (inside vmap, traced once) | R_look.
"""
from __future__ import annotations

import numpy as np
import torch

from . import model as OM
from . import prng as OP

NET = "lookahead_encoder_net"
HEAD = "lookahead_posterior/lookahead_block/linear"


def leaf_shapes(spec: OM.ModelSpec, R: int, H: int):
    out = [(OM._lin_name(NET, 0), 2 * spec.D, H)]
    out += [(OM._lin_name(NET, i), H, H) for i in range(1, 2 * R + 1)]
    out.append((HEAD, H, 2 * spec.d * spec.D))
    return out


def init_params(spec: OM.ModelSpec, R: int, H: int, seed: int = 11, dtype=torch.float64):
    rng = np.random.default_rng(seed)
    p = {}
    for name, fi, fo in leaf_shapes(spec, R, H):
        w = np.clip(rng.standard_normal((fi, fo)), -2, 2) / np.sqrt(fi)
        p[name] = {"w": torch.tensor(w.astype(np.float32), dtype=dtype),
                   "b": torch.tensor((0.05 * rng.standard_normal(fo)).astype(np.float32), dtype=dtype)}
    return p


def lookahead_encoder(lp, spec: OM.ModelSpec, R: int, ln: bool, x_o_b):
    """hk.Sequential([ResidualMLP, LookaheadBlock]) (lookahead.py:14-39,77-80) -> (loc, scale), each [B, F, d]."""
    h = OM.residual_mlp(lp, NET, x_o_b, R, ln)
    par = OM.linear(lp, HEAD, h).reshape(x_o_b.shape[0], spec.D, 2 * spec.d)
    return par[..., :spec.d], torch.nn.functional.softplus(par[..., spec.d:]) + 1e-5


def diag_log_prob(z, loc, scale):
    return (-0.5 * ((z - loc) / scale) ** 2 - torch.log(scale) - 0.5 * OM.LOG2PI).sum(-1)


def model_one_step_samples(p, spec: OM.ModelSpec, x, b, rng_key, K: int, S: int):
    """lookahead.py:123-176 -> (subsampled_inds [S], valid_mask [B,S], model_one_step_z [K,B,S,d]) (no gradients)."""
    B, F, d = x.shape[0], spec.D, spec.d
    dt = x.dtype
    seq = OP.PRNGSequence(rng_key)
    x_o = x * b
    hp = OM.residual_mlp(p, "partial_encoder_net", torch.cat([x_o, b], -1), spec.R_part, spec.ln_part)
    mu_p, L_p = OM.tril_head(p, "partial_posterior_dist/linear", hp, d)
    seq.take(spec.R_part)
    eps = torch.tensor(OP.normal(seq.next(), (K, B, d)), dtype=dt)
    z = mu_p.unsqueeze(0) + torch.einsum("bij,kbj->kbi", L_p, eps)
    loc, _ = OM.decoder(p, spec, z.reshape(K * B, d))
    seq.take(spec.R_dec)
    x_s = torch.where((b == 1).unsqueeze(0), x_o.unsqueeze(0), loc.reshape(K, B, F))
    inds = OP.choice_without_replacement(seq.next(), F, S)
    one_hots = torch.eye(F, dtype=dt)[torch.as_tensor(inds)]
    b_look = torch.maximum(b.unsqueeze(1), one_hots.unsqueeze(0))                      # [B,S,F]
    x_look = x_s.unsqueeze(2) * b_look.unsqueeze(0)                                    # [K,B,S,F]
    valid = ((b.unsqueeze(1) + one_hots.unsqueeze(0)).amax(-1) < 2).to(dt)
    keys = OP.split(seq.next(), K)
    bl = b_look.reshape(B * S, F)
    z1 = []
    for k in range(K):
        h = OM.residual_mlp(p, "partial_encoder_net", torch.cat([x_look[k].reshape(B * S, F), bl], -1), spec.R_part, spec.ln_part)
        mu, L = OM.tril_head(p, "partial_posterior_dist/linear", h, d)
        e = torch.tensor(OP.normal(keys[k], (1, B * S, d))[0], dtype=dt)
        z1.append(mu + torch.einsum("bij,bj->bi", L, e))
    return inds, valid, torch.stack(z1).reshape(K, B, S, d)


def lookahead_lls(lp, spec: OM.ModelSpec, R: int, ln: bool, x, b, inds, valid, z1):
    """lookahead.py:178-202 -> [B] (the reference's array carries a unit axis in front)."""
    loc, scale = lookahead_encoder(lp, spec, R, ln, torch.cat([x * b, b], -1))
    idx = torch.as_tensor(inds)
    lls = diag_log_prob(z1, loc[:, idx].unsqueeze(0), scale[:, idx].unsqueeze(0))      # [K,B,S]
    lls = lls.mean(0) * valid
    denom = (valid != 0).sum(-1)
    out = lls.sum(-1) / denom.clamp(min=1)
    return torch.where(denom == 0, torch.zeros_like(out), out)


def loss_and_grads(p, lp, spec: OM.ModelSpec, R: int, ln: bool, x, b, rng_key, K: int, S: int):
    with torch.no_grad():
        inds, valid, z1 = model_one_step_samples(p, spec, x, b, rng_key, K, S)
    q = {n: {k: t.detach().clone().requires_grad_(True) for k, t in leaf.items()} for n, leaf in lp.items()}
    ll = lookahead_lls(q, spec, R, ln, x, b, inds, valid, z1)
    loss = -ll.mean()
    loss.backward()
    return loss.detach(), ll.detach(), {n: {k: t.grad for k, t in leaf.items()} for n, leaf in q.items()}, (inds, valid, z1)


def expected_info_gains(p, lp, spec: OM.ModelSpec, R: int, ln: bool, x, b):
    """lookahead.py:204-227 for one instance x [F], b [F]."""
    d = spec.d
    x, b = x.reshape(1, -1), b.reshape(1, -1)
    h = OM.residual_mlp(p, "encoder_net", x, spec.R_enc, spec.ln_enc)
    _, L = OM.tril_head(p, "posterior_dist/linear", h, d)
    cur = 0.5 * d * (1.0 + OM.LOG2PI) + torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
    _, scale = lookahead_encoder(lp, spec, R, ln, torch.cat([x * b, b], -1))
    ents = (0.5 * (1.0 + OM.LOG2PI) + torch.log(scale)).sum(-1).reshape(-1)
    gains = cur - ents
    return torch.where(b.reshape(-1) == 0, gains, torch.full_like(gains, -float("inf")))
