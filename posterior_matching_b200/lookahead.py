"""`LookaheadPosterior` (reference: posterior_matching/models/lookahead.py:14-227, train_lookahead_posterior.py:47-70;
SURVEY.md §8f N3) over a feature-vector `PosteriorMatchingVAE` or a frozen convolutional one with a TriLGaussian partial
posterior (`ConvPosteriorMatchingVAE`: configs/pm_vae_mnist16.py + configs/lookahead_mnist16.py).

A frozen PM-VAE produces, for every training row, `model_samples` imputations of the unobserved features and from each
of them one latent sample of q(z | x_o + one more feature) for `lookahead_subsample` candidate features; a second
encoder (`lookahead_encoder_net` + `LookaheadBlock`) learns one diagonal-Gaussian "lookahead posterior" per feature to
match those samples.  The work is almost entirely in the frozen model: K * B * S rows through the partial encoder
(the fused tcgen05 chains behind `pmvae_net_apply`, or the convolution operators of conv_vae.py) and the TriL sampling
kernels; the trained encoder sees B rows.  Its Linears run through `pmvae_linear` / `pmvae_linear_backward`, a
ConvEncoder lookahead net through `pmvae_conv2d_forward / backward`, and the objective through `pmvae_lookahead_ll`
(csrc/dist_ops.cu); torch.autograd only strings those operators together (relu / LayerNorm / adds on [B, H]).

Key order of one call, as Haiku hands them out (`hk.next_rng_key()`; every ResidualMLP block draws a dropout key even at
rate 0, networks.py:124; the convolutional networks draw none) [R: restated from the reference's call order, no JAX here to replay it]:
    R_part keys (partial encoder) | z sample | R_dec keys (decoder) | choice | split -> K sample keys | ...
`jax.random.choice(key, F, (S,), replace=False)` is `permutation(key, F)[:S]`: one `sort_key_val` round per
ceil(3 ln F / ln(2^32 - 1)) with 32 random bits per element from `split(key)[1]` [R: jax 0.2.26 `_shuffle`].
"""
from __future__ import annotations

import math
from typing import Any, Dict, Mapping, Optional

import numpy as np
import torch

from . import _lib, prng
from .conv_vae import ConvPosteriorMatchingVAE, _ConvStack
from .dist_objects import MultivariateNormalTriL
from .vae import PosteriorMatchingVAE, ResidualMLP, _f32c, _stream, get_network


class _LinearFn(torch.autograd.Function):
    """hk.Linear on the library's float32 GEMM kernels, forward and VJP."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous()
        B, K = x.shape
        N = w.shape[1]
        y = torch.empty((B, N), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib.pmvae_linear(_lib.PREC_F32, x.data_ptr(), w.data_ptr(), b.data_ptr(), B, K, N, 0, y.data_ptr(),
                                         None, 0, _stream()), "pmvae_linear")
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        B, K = x.shape
        N = w.shape[1]
        dx = torch.empty_like(x)
        dw = torch.zeros_like(w)
        db = torch.zeros(N, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib.pmvae_linear_backward(x.data_ptr(), w.data_ptr(), dy.data_ptr(), B, K, N, 0, dx.data_ptr(),
                                                  dw.data_ptr(), db.data_ptr(), _stream()), "pmvae_linear_backward")
        return dx, dw, db


class _LookaheadLLFn(torch.autograd.Function):
    """lookahead.py:183-199 on `pmvae_lookahead_ll`: (par [B,S,2d], z [K,B,S,d], valid [B,S]) -> ll [B]."""

    @staticmethod
    def forward(ctx, par, z, valid):
        par, z, valid = par.contiguous(), z.contiguous(), valid.contiguous()
        K, B, S, d = z.shape
        out = torch.empty(B, dtype=torch.float32, device=par.device)
        _lib.check(_lib.lib.pmvae_lookahead_ll(par.data_ptr(), z.data_ptr(), valid.data_ptr(), K, B, S, d, out.data_ptr(),
                                               _stream()), "pmvae_lookahead_ll")
        ctx.save_for_backward(par, z, valid)
        return out

    @staticmethod
    def backward(ctx, g):
        par, z, valid = ctx.saved_tensors
        K, B, S, d = z.shape
        dpar = torch.empty_like(par)
        g = g.contiguous()
        _lib.check(_lib.lib.pmvae_lookahead_ll_backward(par.data_ptr(), z.data_ptr(), valid.data_ptr(), g.data_ptr(), K, B,
                                                        S, d, dpar.data_ptr(), _stream()), "pmvae_lookahead_ll_backward")
        return dpar, None, None


class ConvEncoder:
    """networks.py:9-38 (spec only): `conv_layers` = [(filters, kernel, stride), ...]."""

    def __init__(self, conv_layers, name: Optional[str] = None):
        self.conv_layers, self.name = [tuple(int(v) for v in l) for l in conv_layers], name


class _ConvStackFn(torch.autograd.Function):
    """A ConvEncoder stack on `pmvae_conv2d_forward / backward` (conv_vae._ConvStack), forward and VJP."""

    @staticmethod
    def forward(ctx, x, stack, *flat):
        params = {n: {"w": flat[2 * i], "b": flat[2 * i + 1]} for i, n in enumerate(stack.names)}
        acts = stack.forward(params, x.contiguous())
        ctx.stack, ctx.params, ctx.acts = stack, params, acts
        return acts[-1]

    @staticmethod
    def backward(ctx, dy):
        grads = {n: {k: torch.zeros_like(t) for k, t in leaf.items()} for n, leaf in ctx.params.items()}
        ctx.stack.backward(ctx.params, grads, ctx.acts, dy.contiguous(), need_dx=False)
        out = [None, None]
        for n in ctx.stack.names:
            out += [grads[n]["w"], grads[n]["b"]]
        return tuple(out)


def _layer_norm(h, eps: float = 1e-5):
    """hk.LayerNorm(-1, False, False) (networks.py:118): biased variance, no scale / offset."""
    mu = h.mean(-1, keepdim=True)
    var = ((h - mu) ** 2).mean(-1, keepdim=True)
    return (h - mu) * torch.rsqrt(var + eps)


def _lin_name(prefix: str, i: int) -> str:
    return f"{prefix}/linear" if i == 0 else f"{prefix}/linear_{i}"


class LookaheadBlock:
    """lookahead.py:14-39 (spec): one hk.Linear to 2 * event_size * num_features, read as [B, F, loc | raw scale]."""

    def __init__(self, event_size: int, num_features: int, name: Optional[str] = None):
        self.event_size, self.num_features, self.name = int(event_size), int(num_features), name
        self.num_params = 2 * self.event_size


class LookaheadPosterior:
    NET = "lookahead_encoder_net"
    HEAD = "lookahead_posterior/lookahead_block/linear"        # [R] Haiku module path of LookaheadBlock's Linear

    def __init__(self, pm_vae, lookahead_encoder_net, num_features: int, lookahead_subsample: int = 16,
                 model_samples: int = 64, name: Optional[str] = None):
        if not isinstance(pm_vae, (PosteriorMatchingVAE, ConvPosteriorMatchingVAE)):
            raise TypeError("pm_vae must be a PosteriorMatchingVAE or a ConvPosteriorMatchingVAE")
        if isinstance(pm_vae, ConvPosteriorMatchingVAE) and pm_vae.argmm is not None:
            raise NotImplementedError("LookaheadPosterior needs a TriLGaussian partial posterior (q(z | x_o).sample)")
        if int(num_features) != pm_vae.num_features:
            raise ValueError("num_features must match the PM-VAE's")
        self.pm_vae, self.net, self.name = pm_vae, lookahead_encoder_net, name
        self.device = pm_vae.device
        self.feature_shape = tuple(getattr(pm_vae, "feature_shape", (pm_vae.num_features,)))
        self.block = LookaheadBlock(pm_vae.latent_dim, num_features)
        self._num_features, self._lookahead_subsample, self._model_samples = int(num_features), int(lookahead_subsample), int(model_samples)
        if self._lookahead_subsample > self._num_features:
            raise ValueError("lookahead_subsample exceeds num_features (jax.random.choice without replacement would fail)")
        # leaves: (Haiku name, weight shape, bias size)
        self.stack = None
        if isinstance(self.net, ResidualMLP):
            if len(self.feature_shape) != 1:
                raise NotImplementedError("a ResidualMLP lookahead encoder takes feature vectors")
            H, R = self.net.hidden_units, self.net.residual_blocks
            self.leaves = [(_lin_name(self.NET, 0), (2 * num_features, H), H)]
            self.leaves += [(_lin_name(self.NET, i), (H, H), H) for i in range(1, 2 * R + 1)]
            feat = H
        elif isinstance(self.net, ConvEncoder):
            if len(self.feature_shape) != 3 or self.feature_shape[0] != self.feature_shape[1]:
                raise NotImplementedError("a ConvEncoder lookahead encoder takes square [H, W, C] images")
            self.stack = _ConvStack(self.NET, self.net.conv_layers, self.feature_shape[0], 2 * self.feature_shape[2], False)
            self.leaves = list(self.stack.leaf_shapes())
            feat = self.stack.out_hw ** 2 * self.stack.out_c
        else:
            raise NotImplementedError("lookahead_encoder_net must be a ResidualMLP or a ConvEncoder spec")
        self.leaves.append((self.HEAD, (feat, self.block.num_params * num_features), self.block.num_params * num_features))
        n = sum(int(np.prod(ws)) + nb for _, ws, nb in self.leaves)
        self.arena = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grad_arena = torch.zeros_like(self.arena)
        self.m, self.v = torch.zeros_like(self.arena), torch.zeros_like(self.arena)
        self.params = self._views(self.arena)
        self.grads = self._views(self.grad_arena)
        self.step = 0

    @classmethod
    def from_config(cls, config: Mapping[str, Any], pm_vae_config: Mapping[str, Any], name: Optional[str] = None,
                    **pm_vae_kwargs) -> "LookaheadPosterior":
        """lookahead.py:84-120: the lookahead encoder defaults to the PM-VAE's encoder type and config."""
        pm_vae = PosteriorMatchingVAE.from_config(pm_vae_config, **pm_vae_kwargs)      # conv configs: ConvPosteriorMatchingVAE
        net_type = config.get("lookahead_encoder_net", pm_vae_config["encoder_net"])
        net_cfg = config.get("lookahead_encoder_net_config", pm_vae_config.get("encoder_net_config"))
        if net_type == "ConvEncoder":
            net = ConvEncoder(**dict(net_cfg or {}), name=cls.NET)
        else:
            net = get_network(net_type, net_cfg, name=cls.NET)
        return cls(pm_vae, net, config["num_features"], config.get("lookahead_subsample", 16),
                   config.get("model_samples", 64), name=name)

    # ---- parameters ---------------------------------------------------------------------------------
    def _views(self, arena: torch.Tensor) -> Dict[str, Dict[str, torch.Tensor]]:
        out, off = {}, 0
        for nm, ws, nb in self.leaves:
            nw = int(np.prod(ws))
            out[nm] = {"w": arena[off:off + nw].view(*ws), "b": arena[off + nw:off + nw + nb]}
            off += nw + nb
        return out

    def init(self, seed: int = 0):
        """Haiku defaults: w ~ TruncatedNormal(+-2 sigma) / sqrt(fan_in), b = 0 (the lookahead modules only; the PM-VAE
        is loaded from its own checkpoint, train_lookahead_posterior.py:38-42)."""
        g = torch.Generator(device="cpu").manual_seed(int(seed))
        for leaf in self.params.values():
            w = torch.empty(leaf["w"].shape, dtype=torch.float32)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
            leaf["w"].copy_(w / math.sqrt(w[..., 0].numel()))
            leaf["b"].zero_()
        return self.params

    def load_params(self, params: Mapping[str, Mapping[str, Any]], strict: bool = True):
        for nm, leaf in self.params.items():
            if nm not in params:
                if strict:
                    raise KeyError(nm)
                continue
            for k, dst in leaf.items():
                src = params[nm][k]
                src = src if torch.is_tensor(src) else torch.as_tensor(np.asarray(src))
                dst.copy_(src.to(device=self.device, dtype=torch.float32).reshape(dst.shape))

    # ---- lookahead encoder ----------------------------------------------------------------------------
    def lookahead_encoder(self, x_o_b: torch.Tensor, params=None) -> torch.Tensor:
        """hk.Sequential([lookahead_encoder_net, LookaheadBlock]) (lookahead.py:77-80) -> raw parameters [B, F, 2d]."""
        p = self.params if params is None else params
        x_o_b = _f32c(x_o_b, self.device)
        if self.stack is not None:
            flat = [p[n][k] for n in self.stack.names for k in ("w", "b")]
            h = _ConvStackFn.apply(x_o_b, self.stack, *flat)
            h = h.reshape(h.shape[0], -1)               # LookaheadBlock: rearrange "b ... -> b (...)" (lookahead.py:25)
        else:
            ln, R = self.net.layer_norm, self.net.residual_blocks

            def lin(i, h):
                leaf = p[_lin_name(self.NET, i)]
                h = _LinearFn.apply(h, leaf["w"], leaf["b"])
                return _layer_norm(h) if ln else h

            h = lin(0, x_o_b)
            for r in range(R):
                res = lin(2 * r + 1, torch.relu(h))
                res = lin(2 * r + 2, torch.relu(res))
                h = h + res
            h = torch.relu(h)
        out = _LinearFn.apply(h, p[self.HEAD]["w"], p[self.HEAD]["b"])
        return out.view(out.shape[0], self._num_features, self.block.num_params)

    # ---- jax.random.choice(key, n, (k,), replace=False) -----------------------------------------------
    def _choice(self, key, n: int, k: int) -> torch.Tensor:
        perm = torch.arange(n, device=self.device)
        rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
        for _ in range(rounds):
            key, sub = prng.split(key, 2)
            bits = torch.empty(n, dtype=torch.int32, device=self.device)
            _lib.check(_lib.lib.pmvae_random_bits(_lib.key_arg(sub), n, 0, n, bits.data_ptr(), _stream()), "pmvae_random_bits")
            order = torch.sort(bits.to(torch.int64) & 0xFFFFFFFF, stable=True).indices
            perm = perm[order]
        return perm[:k]

    # ---- lookahead.py:122-202 -------------------------------------------------------------------------
    @torch.no_grad()
    def model_one_step_samples(self, x: torch.Tensor, b: torch.Tensor, rng):
        """The frozen-model half of `__call__`: (subsampled_inds [S], valid_mask [B,S], model_one_step_z [K,B,S,d])."""
        pm, F, K, S = self.pm_vae, self._num_features, self._model_samples, self._lookahead_subsample
        d, shape = pm.latent_dim, self.feature_shape
        x, b = _f32c(x, self.device), _f32c(b, self.device)
        B = x.shape[0]
        seq = prng.PRNGSequence(rng)
        x_o = x * b
        po_posterior = pm.partial_encoder(torch.cat([x_o, b], dim=-1))
        seq.skip(pm.cfg.R_part)
        z = po_posterior.sample(seed=seq.next(), sample_shape=K)                      # [K, B, d]
        dec_mean = pm.decoder(z.view(K * B, d)).mean().reshape(K, B, *shape)
        seq.skip(pm.cfg.R_dec)
        x_samples = torch.where((b == 1).unsqueeze(0), x_o.unsqueeze(0), dec_mean)     # [K, B, *shape]
        inds = self._choice(seq.next(), F, S)
        one_hots = torch.eye(F, device=self.device)[inds].view(S, *shape)              # lookahead.py:139-150
        b_look = torch.maximum(b.unsqueeze(1), one_hots.unsqueeze(0))                  # [B, S, *shape]
        x_look = (x_samples.unsqueeze(2) * b_look.unsqueeze(0)).reshape(K * B * S, *shape)
        valid = ((b.unsqueeze(1) + one_hots.unsqueeze(0)).flatten(2).amax(dim=-1) < 2).to(torch.float32)
        keys = prng.split(seq.next(), K)
        b_rep = b_look.unsqueeze(0).expand(K, *b_look.shape).reshape(K * B * S, *shape)
        par = pm.partial_encoder(torch.cat([x_look, b_rep], dim=-1)).parameters
        par = par.view(K, B * S, -1)
        z1 = torch.empty((K, B * S, d), dtype=torch.float32, device=self.device)
        for k in range(K):                     # jax.vmap(model_sample) over the K keys: one [B*S, d] draw per key
            z1[k] = MultivariateNormalTriL(par[k], d).sample(seed=keys[k])
        return inds, valid, z1.view(K, B, S, d)

    def __call__(self, x: torch.Tensor, b: torch.Tensor, is_training: bool = False, *, rng=None, params=None) -> torch.Tensor:
        """-> lookahead_lls [1, B] (the reference's rearrange "(z b) ... -> z b ..." of a [K, B, S] array leaves a unit
        axis in front, lookahead.py:190-199); differentiable with respect to the lookahead parameters."""
        if rng is None:
            raise ValueError("pass rng= (the key hk.transform's apply would receive)")
        inds, valid, z1 = self.model_one_step_samples(x, b, rng)
        x, b = _f32c(x, self.device), _f32c(b, self.device)
        raw = self.lookahead_encoder(torch.cat([x * b, b], dim=-1), params)
        ll = _LookaheadLLFn.apply(raw[:, inds].contiguous(), z1, valid)
        return ll.view(1, -1)

    def expected_info_gains(self, x: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """lookahead.py:204-227 for one instance x, b of the feature shape -> [F]; -inf where the feature is observed."""
        pm, F, d, shape = self.pm_vae, self._num_features, self.pm_vae.latent_dim, self.feature_shape
        x, b = _f32c(x, self.device).view(1, *shape), _f32c(b, self.device).view(1, *shape)
        with torch.no_grad():
            current_ent = pm.encoder(x).entropy()                                      # [1]
            raw = self.lookahead_encoder(torch.cat([x * b, b], dim=-1)).view(F, 2 * d).contiguous()
            ents = torch.empty(F, dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib.pmvae_diag_log_prob(raw.data_ptr(), None, F, d, None, ents.data_ptr(), _stream()),
                       "pmvae_diag_log_prob")
            gains = current_ent - ents
            return torch.where(b.reshape(F) == 0, gains, torch.full_like(gains, -math.inf))

    # ---- train_lookahead_posterior.py:47-70 -----------------------------------------------------------
    def loss_and_grads(self, x, b, *, rng):
        leaves = {nm: {k: t.detach().requires_grad_(True) for k, t in leaf.items()} for nm, leaf in self.params.items()}
        ll = self(x, b, rng=rng, params=leaves)
        loss = -ll.mean()
        flat = [t for leaf in leaves.values() for t in leaf.values()]
        gs = torch.autograd.grad(loss, flat, allow_unused=True)
        self.grad_arena.zero_()
        it = iter(gs)
        for nm, leaf in self.grads.items():
            for k in leaf:
                g = next(it)
                if g is not None:
                    leaf[k].copy_(g)
        return loss.detach(), self.grads

    def train_step(self, x, b, *, rng, lr_schedule=None, adam=(0.9, 0.999, 1e-8)) -> Dict[str, float]:
        """loss = -mean(lookahead_lls); scale_by_adam -> scale_by_schedule -> scale(-1) on the lookahead modules only
        (`trainable_predicate`, train_lookahead_posterior.py:60-61)."""
        loss, _ = self.loss_and_grads(x, b, rng=rng)
        lr = float(lr_schedule(self.step)) if lr_schedule else 1e-3
        _lib.check(_lib.lib.pmvae_adamw_flat(self.arena.data_ptr(), self.grad_arena.data_ptr(), self.m.data_ptr(),
                                             self.v.data_ptr(), self.arena.numel(), self.step, lr, 0.0, adam[0], adam[1],
                                             adam[2], _stream()), "pmvae_adamw_flat")
        self.step += 1
        return {"loss": float(loss)}
