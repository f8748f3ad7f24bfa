"""The Posterior Matching term of `PosteriorMatchingVADE` (reference: posterior_matching/models/vade.py:150-265,
train_pm_vade.py:38-61) for configs/pm_vade_mnist.py: ConvEncoder + DiagonalGaussian posterior
(distributions.py:58-84), ConvEncoder partial encoder + AutoregressiveGMM partial posterior, composed on the host from
libpmvae operators like conv_vae.py.

    posterior_matching_ll(x, b) = log q(stop_gradient(z) | x_o),   z ~ q(z | x) = N(loc, diag(softplus(raw) + 1e-5))

Only the `partial_*` modules are trained (train_pm_vade.py:57-58 `trainable_predicate`): `backward` produces their
gradients, `train_step` applies the optax chain of train_pm_vade.py:49-53 (Adam -> schedule -> -1, no weight decay)
to them.  The VaDE's own ELBO / clustering heads (vade.py:100-148), trained by train_vade.py, are outside the hot path.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Mapping, Optional

import numpy as np
import torch

from . import _lib, prng
from .conv_vae import _ConvStack, _f32c, _stream
from .distributions import AutoregressiveGMM


class PosteriorMatchingVADE:
    def __init__(self, config: Mapping[str, Any], *, device=None, image_size: int = 28, channels: int = 1,
                 precision: str = "fp32"):
        if not torch.cuda.is_available():
            raise RuntimeError("PosteriorMatchingVADE needs a CUDA device: the hot path has no CPU fallback")
        if config["encoder_net"] != "ConvEncoder" or config.get("partial_posterior_dist", "TriLGaussian") != "AutoregressiveGMM":
            raise NotImplementedError("this class covers the combination configs/pm_vade_mnist.py uses")
        self.device = torch.device("cuda" if device is None else device)
        self.latent_dim = d = int(config["latent_dim"])
        enc_layers = [tuple(l) for l in config["encoder_net_config"]["conv_layers"]]
        part_layers = [tuple(l) for l in config.get("partial_encoder_net_config", config["encoder_net_config"])["conv_layers"]]
        self.enc = _ConvStack("encoder_net", enc_layers, image_size, channels, False, precision)
        self.part = _ConvStack("partial_encoder_net", part_layers, image_size, 2 * channels, False, precision)
        self.enc_feat = self.enc.out_hw ** 2 * self.enc.out_c
        ar = dict(config.get("partial_posterior_dist_config", {}) or {})
        self.argmm = AutoregressiveGMM(d, ar.get("num_components", 10), ar.get("residual_blocks", 2),
                                       ar.get("hidden_units", 256), context_size=self.part.out_hw ** 2 * self.part.out_c,
                                       device=self.device, precision=precision)
        # frozen (pre-trained VaDE) leaves and trainable partial-encoder leaves live in two arenas
        self.frozen_leaves = self.enc.leaf_shapes() + [("diagonal_gaussian/linear", (self.enc_feat, 2 * d), 2 * d)]
        self.train_leaves = self.part.leaf_shapes()
        self.frozen_arena = torch.zeros(sum(int(np.prod(s)) + nb for _, s, nb in self.frozen_leaves), device=self.device)
        self.arena = torch.zeros(sum(int(np.prod(s)) + nb for _, s, nb in self.train_leaves), device=self.device)
        self.grad_arena = torch.zeros_like(self.arena)
        self.params = {**self._views(self.frozen_arena, self.frozen_leaves), **self._views(self.arena, self.train_leaves),
                       **self.argmm.params}
        self.grads = {**self._views(self.grad_arena, self.train_leaves), **self.argmm.grads}
        self.m = [torch.zeros_like(self.arena), torch.zeros_like(self.argmm.arena)]
        self.v = [torch.zeros_like(self.arena), torch.zeros_like(self.argmm.arena)]
        self.step = 0
        self._last = None

    @classmethod
    def from_config(cls, config: Mapping[str, Any], **kw) -> "PosteriorMatchingVADE":
        """vade.py:180-226."""
        return cls(config, **kw)

    @staticmethod
    def _views(arena, leaves):
        out, off = {}, 0
        for name, shape, nb in leaves:
            nw = int(np.prod(shape))
            out[name] = {"w": arena[off:off + nw].view(*shape), "b": arena[off + nw:off + nw + nb]}
            off += nw + nb
        return out

    def init(self, seed: int = 0):
        g = torch.Generator(device="cpu").manual_seed(int(seed))
        for name, leaf in self.params.items():
            w = torch.empty(leaf["w"].shape, dtype=torch.float32)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
            leaf["w"].copy_(w / math.sqrt(w[..., 0].numel()))
            leaf["b"].zero_()
        return self.params

    def load_params(self, params, strict: bool = True):
        """Copies `{module: {leaf: array}}` (e.g. the pre-trained VaDE's TrainState.params, train_pm_vade.py:45-46)."""
        for name, leaf in self.params.items():
            if name not in params:
                if strict:
                    raise KeyError(name)
                continue
            for k, dst in leaf.items():
                src = params[name][k]
                src = src if torch.is_tensor(src) else torch.as_tensor(np.asarray(src))
                dst.copy_(src.to(device=self.device, dtype=torch.float32).reshape(dst.shape))

    # ---- vade.py:246-265 ----------------------------------------------------------------------------
    def posterior_matching_ll(self, x: torch.Tensor, b: torch.Tensor, *, rng=None, eps: Optional[torch.Tensor] = None,
                              row_start: int = 0, total_rows: Optional[int] = None) -> torch.Tensor:
        x, b = _f32c(x, self.device), _f32c(b, self.device)
        B, d, S = x.shape[0], self.latent_dim, _stream()
        if eps is None:
            if rng is None:
                raise ValueError("pass rng= or eps=")
            key = prng.PRNGSequence(rng).next()            # the conv encoder draws no dropout keys
            total = B if total_rows is None else int(total_rows)
            eps = torch.empty((B, d), dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib.pmvae_normal(_lib.key_arg(key), total * d, int(row_start) * d, B * d, eps.data_ptr(), S),
                       "pmvae_normal")
        eps = _f32c(eps, self.device)
        feat = self.enc.forward(self.params, x)[-1].reshape(B, self.enc_feat)
        par = torch.empty((B, 2 * d), dtype=torch.float32, device=self.device)
        hw = self.params["diagonal_gaussian/linear"]
        _lib.check(_lib.lib.pmvae_linear(_lib.PREC_F32, feat.data_ptr(), hw["w"].data_ptr(), hw["b"].data_ptr(), B,
                                         self.enc_feat, 2 * d, 0, par.data_ptr(), None, 0, S), "pmvae_linear")
        z = torch.empty((B, d), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_diag_sample(par.data_ptr(), eps.data_ptr(), B, d, z.data_ptr(), S), "pmvae_diag_sample")
        part_acts = self.part.forward(self.params, torch.cat([x * b, b], dim=-1).contiguous())
        ll = self.argmm.log_prob(z, part_acts[-1].reshape(B, -1))
        self._last = (part_acts, B)
        return ll

    def backward(self, g: torch.Tensor):
        """VJP of the last `posterior_matching_ll` for the cotangent g[B], with respect to the trainable (`partial_*`)
        parameters; z carries a stop_gradient (vade.py:263), so nothing flows into the encoder."""
        if self._last is None:
            raise RuntimeError("backward() needs a preceding posterior_matching_ll()")
        part_acts, B = self._last
        self.grad_arena.zero_()
        _, _, dctx = self.argmm.backward(_f32c(g, self.device))
        self.part.backward(self.params, self.grads, part_acts, dctx.view_as(part_acts[-1]).contiguous(), need_dx=False)
        return self.grads

    def train_step(self, x, b, *, rng=None, eps=None, lr_schedule=None, adam=(0.9, 0.999, 1e-8), grad_sync=None,
                   global_rows: Optional[int] = None, row_start: int = 0) -> Dict[str, float]:
        """loss = -mean(posterior_matching_ll) (train_pm_vade.py:38-41) and one update of the partial modules."""
        ll = self.posterior_matching_ll(x, b, rng=rng, eps=eps, row_start=row_start, total_rows=global_rows)
        B = ll.shape[0]
        self.backward(torch.full((B,), -1.0 / (global_rows or B), device=self.device))
        if grad_sync is not None:
            grad_sync([self.grad_arena, self.argmm.grad_arena])
        lr = float(lr_schedule(self.step)) if lr_schedule else 1e-3
        for arena, grads, m, v in ((self.arena, self.grad_arena, self.m[0], self.v[0]),
                                   (self.argmm.arena, self.argmm.grad_arena, self.m[1], self.v[1])):
            _lib.check(_lib.lib.pmvae_adamw_flat(arena.data_ptr(), grads.data_ptr(), m.data_ptr(), v.data_ptr(),
                                                 arena.numel(), self.step, lr, 0.0, adam[0], adam[1], adam[2], _stream()),
                       "pmvae_adamw_flat")
        self.step += 1
        return {"loss": float(-ll.mean())}
