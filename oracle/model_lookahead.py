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


# ------------------------------------------------------------------------------------------------------------------
# The same model over a convolutional PM-VAE with TriLGaussian posteriors and a Bernoulli decoder (configs/
# pm_vae_mnist16.py + configs/lookahead_mnist16.py): images [B, H, W, C], masks of the same shape, the lookahead
# encoder a ConvEncoder on the channel-concatenated [x_o, b].  Conv nets draw no dropout keys, so the key order is
# z | choice | split(K).
from dataclasses import dataclass
from typing import Sequence, Tuple

from . import conv as OC


@dataclass(frozen=True)
class ConvLookSpec:
    image_size: int
    channels: int
    d: int
    enc_layers: Sequence[Tuple[int, int, int]]
    dec_layers: Sequence[Tuple[int, int, int]]
    look_layers: Sequence[Tuple[int, int, int]]

    @property
    def F(self) -> int:
        return self.image_size * self.image_size * self.channels

    @property
    def P(self) -> int:
        return self.d + self.d * (self.d + 1) // 2


def _cname(prefix, base, i):
    return f"{prefix}/{base}" if i == 0 else f"{prefix}/{base}_{i}"


def _enc_out(layers, size):
    for i, (_, k, s) in enumerate(layers):
        size = (size - k) // s + 1 if i == len(layers) - 1 else -(-size // s)
    return size


def conv_init(spec: ConvLookSpec, seed: int = 5, dtype=torch.float64, head_scale: float = 0.1):
    """(frozen PM-VAE params, lookahead params) with Haiku leaf names [R]."""
    rng = np.random.default_rng(seed)

    def leaf(wshape, nb, fan_in, scale=1.0):
        w = np.clip(rng.standard_normal(wshape), -2, 2) / np.sqrt(fan_in) * scale
        return {"w": torch.tensor(w.astype(np.float32), dtype=dtype),
                "b": torch.tensor((0.05 * rng.standard_normal(nb)).astype(np.float32), dtype=dtype)}

    def enc(prefix, layers, cin):
        out = {}
        for i, (f, k, _) in enumerate(layers):
            out[_cname(prefix, "conv2_d", i)] = leaf((k, k, cin, f), f, k * k * cin)
            cin = f
        return out

    p = {}
    p.update(enc("encoder_net", spec.enc_layers, spec.channels))
    feat = _enc_out(spec.enc_layers, spec.image_size) ** 2 * spec.enc_layers[-1][0]
    p["posterior_dist/linear"] = leaf((feat, spec.P), spec.P, feat, head_scale)
    cin = spec.d
    for i, (f, k, _) in enumerate(spec.dec_layers):
        p[_cname("decoder_net", "conv2_d_transpose", i)] = leaf((k, k, f, cin), f, k * k * cin)
        cin = f
    p.update(enc("partial_encoder_net", spec.enc_layers, 2 * spec.channels))
    p["partial_posterior_dist/linear"] = leaf((feat, spec.P), spec.P, feat, head_scale)
    lp = enc(NET, spec.look_layers, 2 * spec.channels)
    lfeat = _enc_out(spec.look_layers, spec.image_size) ** 2 * spec.look_layers[-1][0]
    lp[HEAD] = leaf((lfeat, 2 * spec.d * spec.F), 2 * spec.d * spec.F, lfeat)
    return p, lp


def _convs(p, prefix, base, n):
    return [(p[_cname(prefix, base, i)]["w"], p[_cname(prefix, base, i)]["b"]) for i in range(n)]


def _tril_of(p, head, feat, d):
    par = feat @ p[head]["w"] + p[head]["b"]
    return par[:, :d], OM.fill_scale_tril(par[:, d:], d)


def conv_partial(p, spec: ConvLookSpec, x_o_b):
    h = OC.conv_encoder(_convs(p, "partial_encoder_net", "conv2_d", len(spec.enc_layers)), x_o_b, spec.enc_layers)
    return _tril_of(p, "partial_posterior_dist/linear", h.reshape(x_o_b.shape[0], -1), spec.d)


def conv_model_one_step_samples(p, spec: ConvLookSpec, x, b, rng_key, K: int, S: int):
    B, F, d = x.shape[0], spec.F, spec.d
    shape = tuple(x.shape[1:])
    dt = x.dtype
    seq = OP.PRNGSequence(rng_key)
    x_o = x * b
    mu_p, L_p = conv_partial(p, spec, torch.cat([x_o, b], -1))
    eps = torch.tensor(OP.normal(seq.next(), (K, B, d)), dtype=dt)
    z = mu_p.unsqueeze(0) + torch.einsum("bij,kbj->kbi", L_p, eps)
    logits = OC.conv_decoder(_convs(p, "decoder_net", "conv2_d_transpose", len(spec.dec_layers)), z.reshape(K * B, d),
                             spec.dec_layers)
    x_s = torch.where((b == 1).unsqueeze(0), x_o.unsqueeze(0), torch.sigmoid(logits).reshape(K, B, *shape))
    inds = OP.choice_without_replacement(seq.next(), F, S)
    one_hots = torch.eye(F, dtype=dt)[torch.as_tensor(inds)].reshape(S, *shape)
    b_look = torch.maximum(b.unsqueeze(1), one_hots.unsqueeze(0))                      # [B,S,*shape]
    x_look = x_s.unsqueeze(2) * b_look.unsqueeze(0)                                    # [K,B,S,*shape]
    valid = ((b.unsqueeze(1) + one_hots.unsqueeze(0)).flatten(2).amax(-1) < 2).to(dt)
    keys = OP.split(seq.next(), K)
    bl = b_look.reshape(B * S, *shape)
    z1 = []
    for k in range(K):
        mu, L = conv_partial(p, spec, torch.cat([x_look[k].reshape(B * S, *shape), bl], -1))
        e = torch.tensor(OP.normal(keys[k], (1, B * S, d))[0], dtype=dt)
        z1.append(mu + torch.einsum("bij,bj->bi", L, e))
    return inds, valid, torch.stack(z1).reshape(K, B, S, d)


def conv_lookahead_encoder(lp, spec: ConvLookSpec, x_o_b):
    h = OC.conv_encoder(_convs(lp, NET, "conv2_d", len(spec.look_layers)), x_o_b, spec.look_layers)
    par = (h.reshape(x_o_b.shape[0], -1) @ lp[HEAD]["w"] + lp[HEAD]["b"]).reshape(x_o_b.shape[0], spec.F, 2 * spec.d)
    return par[..., :spec.d], torch.nn.functional.softplus(par[..., spec.d:]) + 1e-5


def conv_lookahead_lls(lp, spec: ConvLookSpec, x, b, inds, valid, z1):
    loc, scale = conv_lookahead_encoder(lp, spec, torch.cat([x * b, b], -1))
    idx = torch.as_tensor(inds)
    lls = diag_log_prob(z1, loc[:, idx].unsqueeze(0), scale[:, idx].unsqueeze(0)).mean(0) * valid
    denom = (valid != 0).sum(-1)
    out = lls.sum(-1) / denom.clamp(min=1)
    return torch.where(denom == 0, torch.zeros_like(out), out)


def conv_loss_and_grads(p, lp, spec: ConvLookSpec, x, b, rng_key, K: int, S: int):
    with torch.no_grad():
        inds, valid, z1 = conv_model_one_step_samples(p, spec, x, b, rng_key, K, S)
    q = {n: {k: t.detach().clone().requires_grad_(True) for k, t in leaf.items()} for n, leaf in lp.items()}
    ll = conv_lookahead_lls(q, spec, x, b, inds, valid, z1)
    loss = -ll.mean()
    loss.backward()
    return loss.detach(), ll.detach(), {n: {k: t.grad for k, t in leaf.items()} for n, leaf in q.items()}, (inds, valid, z1)


def conv_expected_info_gains(p, lp, spec: ConvLookSpec, x, b):
    d = spec.d
    x, b = x.unsqueeze(0), b.unsqueeze(0)
    h = OC.conv_encoder(_convs(p, "encoder_net", "conv2_d", len(spec.enc_layers)), x, spec.enc_layers)
    _, L = _tril_of(p, "posterior_dist/linear", h.reshape(1, -1), d)
    cur = 0.5 * d * (1.0 + OM.LOG2PI) + torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
    _, scale = conv_lookahead_encoder(lp, spec, torch.cat([x * b, b], -1))
    ents = (0.5 * (1.0 + OM.LOG2PI) + torch.log(scale)).sum(-1).reshape(-1)
    gains = cur - ents
    return torch.where(b.reshape(-1) == 0, gains, torch.full_like(gains, -float("inf")))
