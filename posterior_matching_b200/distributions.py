"""Distribution heads of configs/pm_vae_mnist.py on the device (libpmvae, float32):
`Bernoulli` (distributions.py:20-25) and `AutoregressiveGMM` (distributions.py:192-223 with
`OneDimensionalGMM` :116-134 and `_AutoregressiveDistribution.log_prob` :152-166).

They are stand-alone operators with the reference's constructor arguments; `ConvPosteriorMatchingVAE`
(conv_vae.py, what `PosteriorMatchingVAE.from_config` returns for the MNIST config) composes them with the
convolutional encoder / decoder (networks.py:9-72).  Both are differentiable and parity-tested on their own; the
AR-GMM also samples (distributions.py:168-189), which `impute` / `is_log_prob` of the MNIST config need.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t, device):
    return t.to(device=device, dtype=torch.float32).contiguous()


class Bernoulli:
    """`tfd.Bernoulli(logits)` as the decoder distribution: `log_prob(logits, x)` returns the per-row sum
    the model uses (vae.py:127-128), optionally weighted per element (the `b *` of vae.py:199-201)."""

    def __init__(self, name: Optional[str] = None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("Bernoulli needs a CUDA device: the hot path has no CPU fallback")
        self.name = name
        self.device = torch.device("cuda" if device is None else device)
        self._last = None

    def log_prob(self, logits: torch.Tensor, x: torch.Tensor, weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = logits.shape[0]
        logits = _f32c(logits.reshape(B, -1), self.device)
        x = _f32c(x.reshape(B, -1), self.device)
        w = _f32c(weights.reshape(B, -1), self.device) if weights is not None else None
        out = torch.empty(B, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_bernoulli_ll(logits.data_ptr(), x.data_ptr(), w.data_ptr() if w is not None else None,
                                               B, logits.shape[1], out.data_ptr(), _stream()), "pmvae_bernoulli_ll")
        self._last = (logits, x, w)
        return out

    def backward(self, g: torch.Tensor) -> torch.Tensor:
        """d(sum_r g[r] * log_prob[r]) / d logits of the last `log_prob` call."""
        if self._last is None:
            raise RuntimeError("backward() needs a preceding log_prob()")
        logits, x, w = self._last
        g = _f32c(g, self.device)
        out = torch.empty_like(logits)
        _lib.check(_lib.lib.pmvae_bernoulli_ll_backward(logits.data_ptr(), x.data_ptr(),
                                                        w.data_ptr() if w is not None else None, g.data_ptr(),
                                                        logits.shape[0], logits.shape[1], out.data_ptr(), _stream()),
                   "pmvae_bernoulli_ll_backward")
        return out


class AutoregressiveGMM:
    """`AutoregressiveGMM(event_size, num_components, residual_blocks, hidden_units)` applied to a
    context of `context_size` flattened features (fixed at construction because there is no Haiku tracing
    here to infer it).  `log_prob(z, context)` -> [B]; `backward(g)` -> (parameter gradients, dz, dcontext)."""

    def __init__(self, event_size: int, num_components: int = 10, residual_blocks: int = 2, hidden_units: int = 256,
                 name: Optional[str] = None, *, context_size: int, device=None, precision: str = "fp32"):
        if not torch.cuda.is_available():
            raise RuntimeError("AutoregressiveGMM needs a CUDA device: the hot path has no CPU fallback")
        self.name = name
        self.device = torch.device("cuda" if device is None else device)
        self.cfg = _lib.ArgmmConfig()
        self.cfg.d, self.cfg.n_comp, self.cfg.R, self.cfg.H, self.cfg.C = (int(event_size), int(num_components),
                                                                            int(residual_blocks), int(hidden_units),
                                                                            int(context_size))
        # precision="bf16": the hidden 256 x 256 Linears of the ResidualMLP run on the tcgen05 GEMMs with bf16 operands
        # (fp32 accumulation); the first Linear, the mixture head and all density algebra stay float32
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        self.cfg.reserved[0] = 1 if (precision == "bf16" and int(hidden_units) == 256) else 0
        self._cfgp = C.byref(self.cfg)
        n = int(_lib.lib.pmvae_argmm_param_count(self._cfgp))
        if n == 0:
            _lib.check(1, "pmvae_argmm_param_count")
        cnt = _lib.lib.pmvae_argmm_layout(self._cfgp, None, 0)
        arr = (_lib.Leaf * cnt)()
        _lib.lib.pmvae_argmm_layout(self._cfgp, arr, cnt)
        self.leaves = [(l.name.decode(), l.rows, l.cols, int(l.w_off), int(l.b_off)) for l in arr]
        self.arena = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grad_arena = torch.zeros_like(self.arena)
        self.params = self._views(self.arena)
        self.grads = self._views(self.grad_arena)
        self._ws = None
        self._last = None

    def _views(self, arena) -> Dict[str, Dict[str, torch.Tensor]]:
        out = {}
        for name, rows, cols, w_off, b_off in self.leaves:
            out[name] = {"w": arena[w_off:w_off + rows * cols].view(rows, cols), "b": arena[b_off:b_off + cols]}
        return out

    def init(self, seed: int = 0):
        g = torch.Generator(device="cpu").manual_seed(int(seed))
        for leaf in self.params.values():
            w = torch.empty(leaf["w"].shape, dtype=torch.float32)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
            leaf["w"].copy_(w / math.sqrt(w.shape[0]))
            leaf["b"].zero_()
        return self.params

    def load_params(self, params):
        for name, leaf in self.params.items():
            for k, dst in leaf.items():
                src = params[name][k]
                src = src if torch.is_tensor(src) else torch.as_tensor(np.asarray(src))
                dst.copy_(src.to(device=self.device, dtype=torch.float32).reshape(dst.shape))

    def _workspace(self, B: int) -> torch.Tensor:
        need = int(_lib.lib.pmvae_argmm_workspace_bytes(self._cfgp, B))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def log_prob(self, z: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
        z = _f32c(z, self.device)
        B = z.shape[0]
        context = _f32c(context.reshape(B, -1), self.device)      # hk.Flatten (distributions.py:222)
        if z.shape[1] != self.cfg.d or context.shape[1] != self.cfg.C:
            raise ValueError(f"expected z [B, {self.cfg.d}] and context [B, {self.cfg.C}]")
        out = torch.empty(B, dtype=torch.float32, device=self.device)
        ws = self._workspace(B)
        _lib.check(_lib.lib.pmvae_argmm_log_prob(self._cfgp, self.arena.data_ptr(), z.data_ptr(), context.data_ptr(), B,
                                                 out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                   "pmvae_argmm_log_prob")
        self._last = (z, context, B)
        return out

    def sample(self, context: torch.Tensor, num_samples: int, *, key) -> torch.Tensor:
        """`_AutoregressiveDistribution._sample_n(key, n)` (distributions.py:168-189) -> [n, B, d]: n samples per context
        row, drawn one latent dimension at a time.  Noise contract in include/pmvae.h (pmvae_argmm_sample)."""
        B = context.shape[0]
        context = _f32c(context.reshape(B, -1), self.device)
        if context.shape[1] != self.cfg.C:
            raise ValueError(f"expected context [B, {self.cfg.C}]")
        n = int(num_samples)
        out = torch.empty((n, B, self.cfg.d), dtype=torch.float32, device=self.device)
        need = int(_lib.lib.pmvae_argmm_sample_workspace_bytes(self._cfgp, B, n))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        _lib.check(_lib.lib.pmvae_argmm_sample(self._cfgp, self.arena.data_ptr(), context.data_ptr(), B, n,
                                               _lib.key_arg(key), out.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                               _stream()), "pmvae_argmm_sample")
        return out

    def backward(self, g: torch.Tensor) -> Tuple[Dict[str, Dict[str, torch.Tensor]], torch.Tensor, torch.Tensor]:
        if self._last is None:
            raise RuntimeError("backward() needs a preceding log_prob()")
        z, context, B = self._last
        g = _f32c(g, self.device)
        dz = torch.empty_like(z)
        dctx = torch.empty_like(context)
        ws = self._workspace(B)
        _lib.check(_lib.lib.pmvae_argmm_backward(self._cfgp, self.arena.data_ptr(), z.data_ptr(), context.data_ptr(), B,
                                                 g.data_ptr(), self.grad_arena.data_ptr(), dz.data_ptr(),
                                                 dctx.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                   "pmvae_argmm_backward")
        return self.grads, dz, dctx
