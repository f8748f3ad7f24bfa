"""PyTorch (CPU) restatement of the PM-VAE model, loss, optimizer and evaluators.

Test infrastructure only (see oracle/__init__.py).  float64 by default (the
truth the CUDA path is compared with), float32 on request (the CPU baseline).

Reference files restated (relative to /root/reference):
  posterior_matching/models/vae.py:61-118   from_config  -> `ModelSpec`
  posterior_matching/models/vae.py:120-144  __call__     -> `forward`
  posterior_matching/models/vae.py:146-169  impute       -> `impute`
  posterior_matching/models/vae.py:171-226  is_log_prob  -> `is_log_prob`
  posterior_matching/models/networks.py:111-135  ResidualMLP -> `residual_mlp`
  posterior_matching/models/distributions.py:41-55   IdentityGaussian
  posterior_matching/models/distributions.py:101-113 TriLGaussian
  train_pm_vae.py:28-43   get_beta_schedule     -> `beta_schedule`
  train_pm_vae.py:58-72   loss_fn               -> `loss_fn`
  train_pm_vae.py:74-83   optax chain           -> `adamw_update`, `lr_schedule`
  posterior_matching/utils.py:124-136 cyclical_annealing_schedule
  eval_pm_vae_uci.py:82-94 eval_fn              -> `eval_fn`

Third-party semantics (tfp 0.15 FillScaleTriL / MultivariateNormalTriL / KL,
haiku 0.0.5 Linear / LayerNorm, optax 0.1.0 adam / add_decayed_weights /
exponential_decay / linear_schedule) follow SURVEY.md Appendix A.2-A.4:
[V] closed forms are checked against torch.distributions in
tests/test_oracle_model.py; [R] items are recollection, unverified.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, Mapping, Optional, Tuple

import numpy as np
import torch

LOG2PI = math.log(2.0 * math.pi)
Params = Dict[str, Dict[str, torch.Tensor]]


# --------------------------------------------------------------------------- spec
@dataclass(frozen=True)
class ModelSpec:
    """What `PosteriorMatchingVAE.from_config` (vae.py:61-118) resolves to for
    the ResidualMLP + TriLGaussian + IdentityGaussian family (the four UCI
    configs as written: the `masked_posterior_*` keys are never read, F4)."""
    D: int
    d: int
    H: int
    R_enc: int
    R_dec: int
    R_part: int
    ln_enc: bool
    ln_dec: bool
    ln_part: bool
    stop_grad: bool

    @property
    def P(self) -> int:
        return self.d + self.d * (self.d + 1) // 2


def spec_from_config(model_config: Mapping[str, Any]) -> ModelSpec:
    c = model_config
    assert c["encoder_net"] == "ResidualMLP" and c["decoder_net"] == "ResidualMLP"
    assert c.get("partial_encoder_net", c["encoder_net"]) == "ResidualMLP"
    assert c["posterior_dist"] == "TriLGaussian"
    assert c.get("partial_posterior_dist", c["posterior_dist"]) == "TriLGaussian"
    assert c["decoder_dist"] == "IdentityGaussian"
    enc = dict(c.get("encoder_net_config") or {})
    dec = dict(c.get("decoder_net_config") or {})
    part = dict(c.get("partial_encoder_net_config", c.get("encoder_net_config")) or {})
    H = enc.get("hidden_units", 256)
    assert dec.get("hidden_units", 256) == H and part.get("hidden_units", 256) == H
    for n in (enc, dec, part):
        assert n.get("dropout", 0.0) == 0.0, "dropout>0 is out of scope (SURVEY §2)"
    return ModelSpec(
        D=int(c["decoder_dist_config"]["event_size"]), d=int(c["latent_dim"]), H=int(H),
        R_enc=int(enc.get("residual_blocks", 2)), R_dec=int(dec.get("residual_blocks", 2)),
        R_part=int(part.get("residual_blocks", 2)),
        ln_enc=bool(enc.get("layer_norm", False)), ln_dec=bool(dec.get("layer_norm", False)),
        ln_part=bool(part.get("layer_norm", False)),
        stop_grad=bool(c.get("matching_ll_stop_gradients", False)))


def _lin_name(prefix: str, i: int) -> str:
    return f"{prefix}/linear" if i == 0 else f"{prefix}/linear_{i}"


def leaf_shapes(spec: ModelSpec):
    """Haiku parameter leaves in creation order [R naming]."""
    out = []

    def mlp(prefix, fan_in, R):
        out.append((_lin_name(prefix, 0), fan_in, spec.H))
        for i in range(1, 2 * R + 1):
            out.append((_lin_name(prefix, i), spec.H, spec.H))

    mlp("encoder_net", spec.D, spec.R_enc)
    out.append(("posterior_dist/linear", spec.H, spec.P))
    mlp("decoder_net", spec.d, spec.R_dec)
    out.append(("decoder_dist/linear", spec.H, spec.D))
    mlp("partial_encoder_net", 2 * spec.D, spec.R_part)
    out.append(("partial_posterior_dist/linear", spec.H, spec.P))
    return out


def init_params(spec: ModelSpec, seed: int = 3, dtype=torch.float64) -> Params:
    """Haiku default init [R]: w ~ TruncatedNormal(+-2 sigma)*1/sqrt(fan_in), b = 0,
    log_scale = 0 (distributions.py:50-52)."""
    rng = np.random.default_rng(seed)
    p: Params = {}
    for name, fi, fo in leaf_shapes(spec):
        w = np.empty((fi, fo))
        flat = w.reshape(-1)
        n = 0
        while n < flat.size:
            cand = rng.standard_normal(flat.size - n)
            cand = cand[np.abs(cand) <= 2.0]
            flat[n:n + cand.size] = cand
            n += cand.size
        w *= 1.0 / math.sqrt(fi)
        p[name] = {"w": torch.tensor(w.astype(np.float32), dtype=dtype),
                   "b": torch.zeros(fo, dtype=dtype)}
    p["decoder_dist"] = {"log_scale": torch.zeros((), dtype=dtype)}
    return p


def n_params(spec: ModelSpec) -> int:
    return sum(fi * fo + fo for _, fi, fo in leaf_shapes(spec)) + 1


# --------------------------------------------------------------------------- layers
def linear(p: Params, name: str, x):
    return x @ p[name]["w"] + p[name]["b"]


def layer_norm(x, eps: float = 1e-5):
    """hk.LayerNorm(-1, False, False) [R]: biased variance, no affine."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps)


def residual_mlp(p: Params, prefix: str, x, R: int, ln: bool):
    """networks.py:111-135 with dropout 0 and activate_final=True."""
    h = linear(p, _lin_name(prefix, 0), x)
    if ln:
        h = layer_norm(h)
    for r in range(R):
        res = torch.relu(h)
        res = linear(p, _lin_name(prefix, 2 * r + 1), res)
        if ln:
            res = layer_norm(res)
        res = torch.relu(res)
        res = linear(p, _lin_name(prefix, 2 * r + 2), res)
        if ln:
            res = layer_norm(res)
        h = h + res
    return torch.relu(h)


def fill_triangular_index(d: int) -> np.ndarray:
    """[V: TFP doc example] index map: L[i][j] = v[idx[i, j]] for j <= i, else -1."""
    m = d * (d + 1) // 2
    idx = -np.ones((d, d), dtype=np.int64)
    for i in range(d):
        for j in range(i + 1):
            k = i * d + j
            idx[i, j] = d + k if k < m - d else m - 1 - (k - (m - d))
    return idx


def fill_scale_tril(v, d: int):
    """tfb.FillScaleTriL() [V index map; R softplus+1e-5 defaults]."""
    idx = fill_triangular_index(d)
    gather = torch.as_tensor(np.where(idx < 0, 0, idx))
    L = v[..., gather] * torch.as_tensor((idx >= 0), dtype=v.dtype)
    diag = torch.nn.functional.softplus(torch.diagonal(L, dim1=-2, dim2=-1)) + 1e-5
    return torch.tril(L, -1) + torch.diag_embed(diag)


def tril_head(p: Params, name: str, h, d: int):
    """distributions.py:101-113: Linear(H->P); loc, FillScaleTriL."""
    params = linear(p, name, h)
    return params[..., :d], fill_scale_tril(params[..., d:], d)


def tril_log_prob(z, mu, L):
    """[V vs torch] MultivariateNormalTriL.log_prob."""
    d = mu.shape[-1]
    r = torch.linalg.solve_triangular(L, (z - mu).unsqueeze(-1), upper=False).squeeze(-1)
    return (-0.5 * (r ** 2).sum(-1) - torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
            - 0.5 * d * LOG2PI)


def tril_kl_std_normal(mu, L):
    """[V vs torch] KL(N(mu, LL^T) || N(0, I)) (vae.py:130)."""
    d = mu.shape[-1]
    return (-torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
            + 0.5 * (-d + (L ** 2).sum((-2, -1)) + (mu ** 2).sum(-1)))


def net_head(p: Params, spec: ModelSpec, net: str, head: str, x):
    """One ResidualMLP followed by its distribution head's hk.Linear: the raw parameters behind
    `model.encoder(x)`, `model.decoder(z)`, `model.partial_encoder(x_o_b)` (vae.py:47-53)."""
    R, ln = {"encoder_net": (spec.R_enc, spec.ln_enc), "decoder_net": (spec.R_dec, spec.ln_dec),
             "partial_encoder_net": (spec.R_part, spec.ln_part)}[net]
    return linear(p, head, residual_mlp(p, net, x, R, ln))


def std_normal_log_prob(z):
    return -0.5 * (z ** 2).sum(-1) - 0.5 * z.shape[-1] * LOG2PI


def decoder(p: Params, spec: ModelSpec, z):
    """decoder_net + IdentityGaussian (distributions.py:41-55): (loc, log_scale)."""
    h = residual_mlp(p, "decoder_net", z, spec.R_dec, spec.ln_dec)
    return linear(p, "decoder_dist/linear", h), p["decoder_dist"]["log_scale"]


def normal_log_prob(x, loc, log_scale):
    return -0.5 * ((x - loc) * torch.exp(-log_scale)) ** 2 - log_scale - 0.5 * LOG2PI


# --------------------------------------------------------------------------- model
def forward(p: Params, spec: ModelSpec, x, b, eps) -> Dict[str, torch.Tensor]:
    """PosteriorMatchingVAE.__call__ (vae.py:120-144).  eps [B,d] is the N(0,I)
    draw behind `posterior.sample` (z = mu + L eps) [R: key passed unsalted]."""
    h = residual_mlp(p, "encoder_net", x, spec.R_enc, spec.ln_enc)
    mu, L = tril_head(p, "posterior_dist/linear", h, spec.d)
    z = mu + (L @ eps.unsqueeze(-1)).squeeze(-1)
    loc, ls = decoder(p, spec, z)
    rec = normal_log_prob(x, loc, ls).sum(-1)
    kl = tril_kl_std_normal(mu, L)
    x_o_b = torch.cat([x * b, b], -1)
    hp = residual_mlp(p, "partial_encoder_net", x_o_b, spec.R_part, spec.ln_part)
    mu_p, L_p = tril_head(p, "partial_posterior_dist/linear", hp, spec.d)
    zz = z.detach() if spec.stop_grad else z
    match = tril_log_prob(zz, mu_p, L_p)
    return {"reconstruction_ll": rec, "kl": kl, "matching_ll": match, "z": z}


def loss_fn(p: Params, spec: ModelSpec, x, b, eps, beta: float, coef: float = 1.0):
    """train_pm_vae.py:58-72.  Returns (loss, aux-of-means)."""
    out = forward(p, spec, x, b, eps)
    elbo = (out["reconstruction_ll"] - beta * out["kl"]).mean()
    matching_loss = -out["matching_ll"].mean()
    loss = -elbo + coef * matching_loss
    aux = {k: out[k].mean() for k in ("reconstruction_ll", "kl", "matching_ll")}
    aux["beta"] = beta
    return loss, aux


def loss_and_grads(p: Params, spec: ModelSpec, x, b, eps, beta, coef=1.0):
    leaves = [(n, k) for n in p for k in p[n]]
    q = {n: {} for n in p}
    for n, k in leaves:
        q[n][k] = p[n][k].detach().clone().requires_grad_(True)
    loss, aux = loss_fn(q, spec, x, b, eps, beta, coef)
    gs = torch.autograd.grad(loss, [q[n][k] for n, k in leaves], allow_unused=True)
    grads = {n: {} for n in p}
    for (n, k), g in zip(leaves, gs):
        grads[n][k] = torch.zeros_like(p[n][k]) if g is None else g
    return loss.detach(), {k: (v.detach() if torch.is_tensor(v) else v) for k, v in aux.items()}, grads


def impute(p: Params, spec: ModelSpec, x_o, b, eps):
    """vae.py:146-169.  eps [K,B,d] -> imputations [K,B,D]."""
    x_o = x_o * b
    hp = residual_mlp(p, "partial_encoder_net", torch.cat([x_o, b], -1), spec.R_part, spec.ln_part)
    mu_p, L_p = tril_head(p, "partial_posterior_dist/linear", hp, spec.d)
    z = mu_p.unsqueeze(0) + torch.einsum("bij,kbj->kbi", L_p, eps)
    K, B, _ = z.shape
    loc, _ = decoder(p, spec, z.reshape(K * B, -1))
    loc = loc.reshape(K, B, -1)
    return torch.where(b.unsqueeze(0) != 0, x_o.unsqueeze(0), loc)


def is_log_prob(p: Params, spec: ModelSpec, x, b, eps_z, eps_zxo):
    """vae.py:171-226.  eps_* [K,B,d] -> (log p(x), log p(x_u | x_o)), each [B]."""
    h = residual_mlp(p, "encoder_net", x, spec.R_enc, spec.ln_enc)
    mu, L = tril_head(p, "posterior_dist/linear", h, spec.d)
    hp = residual_mlp(p, "partial_encoder_net", torch.cat([x * b, b], -1), spec.R_part, spec.ln_part)
    mu_p, L_p = tril_head(p, "partial_posterior_dist/linear", hp, spec.d)
    K, B, d = eps_z.shape
    z = mu.unsqueeze(0) + torch.einsum("bij,kbj->kbi", L, eps_z)
    z_xo = mu_p.unsqueeze(0) + torch.einsum("bij,kbj->kbi", L_p, eps_zxo)

    def dec_ll(zz, weight):
        loc, ls = decoder(p, spec, zz.reshape(K * B, d))
        ll = normal_log_prob(x.unsqueeze(0), loc.reshape(K, B, -1), ls)
        if weight is not None:
            ll = ll * weight.unsqueeze(0)
        return ll.sum(-1)

    log_p_z = std_normal_log_prob(z)
    log_p_z_xo = std_normal_log_prob(z_xo)
    log_p_xgz = dec_ll(z, None)
    log_q_zgx = tril_log_prob(z, mu.unsqueeze(0), L.unsqueeze(0))
    log_p_xogz = dec_ll(z_xo, b)
    log_q_zgxo = tril_log_prob(z_xo, mu_p.unsqueeze(0), L_p.unsqueeze(0))
    lk = math.log(K)
    log_p_x = torch.logsumexp(log_p_xgz + log_p_z - log_q_zgx, 0) - lk
    log_p_xo = torch.logsumexp(log_p_xogz + log_p_z_xo - log_q_zgxo, 0) - lk
    return log_p_x, log_p_x - log_p_xo


def eval_fn(p: Params, spec: ModelSpec, x, b, eps_imp, eps_z, eps_zxo):
    """eval_pm_vae_uci.py:82-94: (mean-over-K imputation [B,D], log p(x_u|x_o) [B])."""
    imputed = impute(p, spec, x, b, eps_imp).mean(0)
    _, ll = is_log_prob(p, spec, x, b, eps_z, eps_zxo)
    return imputed, ll


# --------------------------------------------------------------------------- schedules / optimizer
def beta_schedule(beta_cfg: Optional[Mapping[str, Any]]):
    """train_pm_vae.py:28-43 + utils.py:124-136 + optax.linear_schedule [R]."""
    if not beta_cfg or "schedule" not in beta_cfg:
        return lambda step: 1.0
    c = beta_cfg
    if c["schedule"] == "monotonic":
        lo, hi, T, begin = c["low_value"], c["high_value"], c["transition_steps"], c["transition_begin"]

        def mono(step):
            frac = min(max((step - begin) / T, 0.0), 1.0)
            return lo + (hi - lo) * frac
        return mono
    if c["schedule"] == "cyclic":
        lo, hi, period, delay = c["low_value"], c["high_value"], c["period"], c.get("delay", 0)

        def cyc(step):
            cnt = step - delay
            cnt = min(max(cnt % period, 0), period // 2)
            frac = 1 - cnt / (period // 2)
            x = (lo - hi) * frac + hi
            return x * (1.0 if step >= delay else 0.0)
        return cyc
    raise ValueError(c["schedule"])


def lr_schedule(init_value, decay_rate, transition_steps):
    """optax.exponential_decay, non-staircase [R]; count starts at 0."""
    return lambda count: init_value * decay_rate ** (count / transition_steps)


def adamw_update(p: Params, g: Params, m: Params, v: Params, count: int, lr: float,
                 wd: float, b1=0.9, b2=0.999, eps=1e-8):
    """train_pm_vae.py:74-83 [R optax 0.1.0]: scale_by_adam -> add_decayed_weights on
    leaves with ndim != 1 (weights AND the 0-d log_scale, F9) -> *lr -> *(-1).
    `count` is the number of updates already applied (0 for the first step);
    bias correction uses t = count + 1, the lr schedule is evaluated by the caller
    at `count`.  In place."""
    t = count + 1
    for n in p:
        for k in p[n]:
            m[n][k].mul_(b1).add_(g[n][k], alpha=1 - b1)
            v[n][k].mul_(b2).addcmul_(g[n][k], g[n][k], value=1 - b2)
            mh = m[n][k] / (1 - b1 ** t)
            vh = v[n][k] / (1 - b2 ** t)
            u = mh / (torch.sqrt(vh) + eps)
            if p[n][k].ndim != 1:
                u = u + wd * p[n][k]
            p[n][k].add_(u, alpha=-lr)


def zeros_like_params(p: Params) -> Params:
    return {n: {k: torch.zeros_like(t) for k, t in d.items()} for n, d in p.items()}


def cast_params(p: Params, dtype) -> Params:
    return {n: {k: t.to(dtype) for k, t in d.items()} for n, d in p.items()}


# --------------------------------------------------------------------------- key chains
def train_eps_key(rng_key, R_enc: int):
    """Sub-key `hk.next_rng_key()` hands to posterior.sample in __call__: the
    encoder's R_enc dropout draws (networks.py:126, F8) come first [R]."""
    from .prng import PRNGSequence
    seq = PRNGSequence(rng_key)
    for _ in range(R_enc):
        seq.next()
    return seq.next()


def eval_keys(rng_key, spec: ModelSpec):
    """Keys behind (impute z, is_log_prob z, is_log_prob z_xo) inside eval_fn
     [R]: partial-enc dropouts, impute z, decoder dropouts (traced
    once under vmap), enc dropouts, partial-enc dropouts, z, z_xo."""
    from .prng import PRNGSequence
    seq = PRNGSequence(rng_key)
    for _ in range(spec.R_part):
        seq.next()
    k_imp = seq.next()
    for _ in range(spec.R_dec + spec.R_enc + spec.R_part):
        seq.next()
    k_z = seq.next()
    k_zxo = seq.next()
    return k_imp, k_z, k_zxo
