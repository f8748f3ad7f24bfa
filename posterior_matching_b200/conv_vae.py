"""PosteriorMatchingVAE for configs/pm_vae_mnist.py and configs/pm_vae_mnist16.py, composed on the host from libpmvae operators
(reference: posterior_matching/models/vae.py:34-144 with ConvEncoder / ConvDecoder networks.py:9-72,
TriLGaussian, Bernoulli and AutoregressiveGMM heads).

SURVEY.md §8f row N1, first cut: float32, correctness-first (direct convolution kernels, fp32 GEMMs);
forward (`__call__`) and `backward` (VJP with per-row cotangents) plus `train_step` (loss_fn of
train_pm_vae.py:58-72 and the optax chain :74-83 with weight_decay = 0, as the MNIST config has).
`impute` / `is_log_prob` (vae.py:146-226) sample the AutoregressiveGMM partial posterior (distributions.py:168-189,
SURVEY §8f N2).
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Any, Dict, Mapping, Optional

import numpy as np
import torch

from . import _lib, prng
from .conv import conv2d_backward, conv2d_forward, conv_desc
from .dist_objects import BernoulliLogits, MultivariateNormalTriL
from .distributions import AutoregressiveGMM, Bernoulli


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t, device):
    return t.to(device=device, dtype=torch.float32).contiguous()


class _ConvStack:
    """ConvEncoder (transpose=False) or ConvDecoder (transpose=True): descriptors + parameter views."""

    def __init__(self, prefix: str, layers, H: int, cin: int, transpose: bool, precision: str = "fp32"):
        self.prefix, self.transpose = prefix, transpose
        self.descs, self.names = [], []
        base = "conv2_d_transpose" if transpose else "conv2_d"
        for i, (f, k, s) in enumerate(layers):
            if transpose:
                pad = "VALID" if i == 0 else "SAME"
            else:
                pad = "VALID" if i == len(layers) - 1 else "SAME"
            d = conv_desc(H, H, cin, f, k, s, pad, transpose=transpose, precision=precision)
            self.descs.append(d)
            self.names.append(f"{prefix}/{base}" if i == 0 else f"{prefix}/{base}_{i}")
            H, cin = d.OH, f
        self.out_hw, self.out_c = H, cin

    def leaf_shapes(self):
        out = []
        for d, n in zip(self.descs, self.names):
            shape = (d.KH, d.KW, d.Cout, d.Cin) if self.transpose else (d.KH, d.KW, d.Cin, d.Cout)
            out.append((n, shape, d.Cout))
        return out

    def forward(self, params, x):
        acts = [x]
        for d, n in zip(self.descs, self.names):
            acts.append(conv2d_forward(d, acts[-1], params[n]["w"], params[n]["b"]))
        return acts

    def backward(self, params, grads, acts, dy, need_dx: bool):
        for i in reversed(range(len(self.descs))):
            d, n = self.descs[i], self.names[i]
            dy = conv2d_backward(d, acts[i], params[n]["w"], acts[i + 1], dy, grads[n]["w"], grads[n]["b"],
                                 need_dx=(i > 0 or need_dx))
        return dy


class ConvPosteriorMatchingVAE:
    def __init__(self, config: Mapping[str, Any], name: Optional[str] = None, *, device=None, image_size: int = 28,
                 channels: int = 1, precision: str = "fp32"):
        if not torch.cuda.is_available():
            raise RuntimeError("ConvPosteriorMatchingVAE needs a CUDA device: the hot path has no CPU fallback")
        if (config["encoder_net"], config["decoder_net"], config["posterior_dist"], config["decoder_dist"]) != (
                "ConvEncoder", "ConvDecoder", "TriLGaussian", "Bernoulli"):
            raise NotImplementedError("this class covers ConvEncoder / ConvDecoder / TriLGaussian / Bernoulli models")
        # vae.py:97-105: the partial posterior defaults to the posterior's type: AutoregressiveGMM in configs/pm_vae_mnist.py,
        # TriLGaussian in configs/pm_vae_mnist16.py (the model LookaheadPosterior is trained over, lookahead_mnist16.py)
        self.partial_posterior_dist = config.get("partial_posterior_dist", config["posterior_dist"])
        if self.partial_posterior_dist not in ("AutoregressiveGMM", "TriLGaussian"):
            raise NotImplementedError(f"partial_posterior_dist {self.partial_posterior_dist!r} is not built")
        self.name = name
        self.device = torch.device("cuda" if device is None else device)
        self.latent_dim = d = int(config["latent_dim"])
        self._stop = bool(config.get("matching_ll_stop_gradients", False))
        enc_layers = [tuple(l) for l in config["encoder_net_config"]["conv_layers"]]
        dec_layers = [tuple(l) for l in config["decoder_net_config"]["conv_layers"]]
        part_layers = [tuple(l) for l in config.get("partial_encoder_net_config", config["encoder_net_config"])["conv_layers"]]
        self.precision = precision       # convolution GEMMs: "fp32" (exact-parity path) or "bf16" (tcgen05, fp32 accumulate)
        self.enc = _ConvStack("encoder_net", enc_layers, image_size, channels, False, precision)
        self.dec = _ConvStack("decoder_net", dec_layers, 1, d, True, precision)
        self.part = _ConvStack("partial_encoder_net", part_layers, image_size, 2 * channels, False, precision)
        if self.dec.out_hw != image_size or self.dec.out_c != channels:
            raise ValueError("decoder does not reproduce the image shape")
        self.P = d + d * (d + 1) // 2
        self.enc_feat = self.enc.out_hw ** 2 * self.enc.out_c
        self.part_feat = self.part.out_hw ** 2 * self.part.out_c
        self.image_size, self.channels = int(image_size), int(channels)
        self.feature_shape = (self.image_size, self.image_size, self.channels)
        self.num_features = self.image_size * self.image_size * self.channels
        self.cfg = SimpleNamespace(R_enc=0, R_dec=0, R_part=0)     # dropout keys a network call draws: none (networks.py:9-72)
        self.argmm = None
        if self.partial_posterior_dist == "AutoregressiveGMM":
            ar_cfg = dict(config.get("partial_posterior_dist_config", {}) or {})
            self.argmm = AutoregressiveGMM(d, ar_cfg.get("num_components", 10), ar_cfg.get("residual_blocks", 2),
                                           ar_cfg.get("hidden_units", 256), context_size=self.part_feat,
                                           device=self.device, precision=precision)
        self.bern = Bernoulli(device=self.device)
        # one flat arena for the conv / head leaves (the AR-GMM keeps its own arena inside `self.argmm`)
        self.leaves = self.enc.leaf_shapes() + [("posterior_dist/linear", (self.enc_feat, self.P), self.P)] + \
            self.dec.leaf_shapes() + self.part.leaf_shapes()
        if self.argmm is None:
            self.leaves.append(("partial_posterior_dist/linear", (self.part_feat, self.P), self.P))
        n = sum(int(np.prod(s)) + nb for _, s, nb in self.leaves)
        self.arena = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grad_arena = torch.zeros_like(self.arena)
        self.params, self.grads = self._views(self.arena), self._views(self.grad_arena)
        self.m, self.v = [torch.zeros_like(self.arena)], [torch.zeros_like(self.arena)]
        if self.argmm is not None:
            self.params.update(self.argmm.params)
            self.grads.update(self.argmm.grads)
            self.m.append(torch.zeros_like(self.argmm.arena))
            self.v.append(torch.zeros_like(self.argmm.arena))
        self.step = 0
        self._last = None

    @classmethod
    def from_config(cls, config: Mapping[str, Any], name: Optional[str] = None, **kw):
        return cls(config, name=name, **kw)

    def _views(self, arena):
        out, off = {}, 0
        for name, shape, nb in self.leaves:
            nw = int(np.prod(shape))
            out[name] = {"w": arena[off:off + nw].view(*shape), "b": arena[off + nw:off + nw + nb]}
            off += nw + nb
        return out

    def load_params(self, params):
        for name, leaf in self.params.items():
            for k, dst in leaf.items():
                src = params[name][k]
                src = src if torch.is_tensor(src) else torch.as_tensor(np.asarray(src))
                dst.copy_(src.to(device=self.device, dtype=torch.float32).reshape(dst.shape))

    def init(self, seed: int = 0):
        g = torch.Generator(device="cpu").manual_seed(int(seed))
        for name, leaf in self.params.items():
            w = torch.empty(leaf["w"].shape, dtype=torch.float32)
            fan_in = w[..., 0].numel() if "transpose" not in name else w.shape[0] * w.shape[1] * w.shape[3]
            torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
            leaf["w"].copy_(w / math.sqrt(fan_in))
            leaf["b"].zero_()
        return self.params

    # ---- vae.py:120-144 ---------------------------------------------------------------------------
    def __call__(self, x: torch.Tensor, b: torch.Tensor, is_training: bool = False, *, rng=None,
                 eps: Optional[torch.Tensor] = None, row_start: int = 0,
                 total_rows: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """`row_start` / `total_rows`: this call's rows inside a global batch (data parallelism): eps is rows
        [row_start, row_start + B) of normal(key, [total_rows, d]), like PosteriorMatchingVAE.draw_eps."""
        x, b = _f32c(x, self.device), _f32c(b, self.device)
        B, d = x.shape[0], self.latent_dim
        if eps is None:
            if rng is None:
                raise ValueError("pass rng= or eps=")
            key = prng.PRNGSequence(rng).next()        # the conv encoder draws no dropout keys
            total = B if total_rows is None else int(total_rows)
            eps = torch.empty((B, d), dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib.pmvae_normal(_lib.key_arg(key), total * d, int(row_start) * d, B * d, eps.data_ptr(),
                                             _stream()), "pmvae_normal")
        eps = _f32c(eps, self.device)
        S = _stream()
        enc_acts = self.enc.forward(self.params, x)
        feat = enc_acts[-1].reshape(B, self.enc_feat)
        par = torch.empty((B, self.P), dtype=torch.float32, device=self.device)
        hw = self.params["posterior_dist/linear"]
        _lib.check(_lib.lib.pmvae_linear(_lib.PREC_F32, feat.data_ptr(), hw["w"].data_ptr(), hw["b"].data_ptr(), B,
                                         self.enc_feat, self.P, 0, par.data_ptr(), None, 0, S), "pmvae_linear")
        z = torch.empty((B, d), dtype=torch.float32, device=self.device)
        kl = torch.empty(B, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_tril_sample_kl(par.data_ptr(), eps.data_ptr(), B, d, z.data_ptr(), kl.data_ptr(), S),
                   "pmvae_tril_sample_kl")
        dec_acts = self.dec.forward(self.params, z.view(B, 1, 1, d))
        rec = self.bern.log_prob(dec_acts[-1], x)
        xob = torch.cat([x * b, b], dim=-1).contiguous()
        part_acts = self.part.forward(self.params, xob)
        ctx = part_acts[-1].reshape(B, -1)
        par_p = None
        if self.argmm is not None:
            match = self.argmm.log_prob(z, ctx)
        else:
            par_p = self._linear(ctx, "partial_posterior_dist/linear", self.part_feat)
            match = torch.empty(B, dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib.pmvae_tril_log_prob(par_p.data_ptr(), z.data_ptr(), B, d, match.data_ptr(), S),
                       "pmvae_tril_log_prob")
        self._last = (x, eps, feat, par, z, enc_acts, dec_acts, part_acts, B, ctx, par_p)
        return {"reconstruction_ll": rec, "kl": kl, "matching_ll": match}

    def _linear(self, h: torch.Tensor, leaf: str, fan_in: int) -> torch.Tensor:
        """One hk.Linear head on [B, fan_in] features -> [B, P] raw TriLGaussian parameters."""
        B = h.shape[0]
        out = torch.empty((B, self.P), dtype=torch.float32, device=self.device)
        hw = self.params[leaf]
        _lib.check(_lib.lib.pmvae_linear(_lib.PREC_F32, h.data_ptr(), hw["w"].data_ptr(), hw["b"].data_ptr(), B, fan_in,
                                         self.P, 0, out.data_ptr(), None, 0, _stream()), "pmvae_linear")
        return out

    def backward(self, g_rec: torch.Tensor, g_kl: torch.Tensor, g_match: torch.Tensor):
        """VJP of the last __call__ for per-row cotangents -> `self.grads` (overwritten)."""
        if self._last is None:
            raise RuntimeError("backward() needs a preceding __call__")
        x, eps, feat, par, z, enc_acts, dec_acts, part_acts, B, ctx, par_p = self._last
        d, S = self.latent_dim, _stream()
        self.grad_arena.zero_()
        g_rec, g_kl, g_match = (_f32c(t, self.device) for t in (g_rec, g_kl, g_match))
        # partial posterior: parameter grads, dz, dcontext -> partial encoder
        if self.argmm is not None:
            _, dz_match, dctx = self.argmm.backward(g_match)
        else:
            dpar_p, dz_match = torch.empty_like(par_p), torch.empty_like(z)
            ws = torch.empty(B * (self.P + d + 1), dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib.pmvae_tril_log_prob_backward(par_p.data_ptr(), z.data_ptr(), g_match.data_ptr(), B, d,
                                                             dpar_p.data_ptr(), dz_match.data_ptr(), ws.data_ptr(),
                                                             ws.numel() * 4, S), "pmvae_tril_log_prob_backward")
            hw, hg = self.params["partial_posterior_dist/linear"], self.grads["partial_posterior_dist/linear"]
            dctx = torch.empty_like(ctx)
            _lib.check(_lib.lib.pmvae_linear_backward(ctx.contiguous().data_ptr(), hw["w"].data_ptr(), dpar_p.data_ptr(), B,
                                                      self.part_feat, self.P, 0, dctx.data_ptr(), hg["w"].data_ptr(),
                                                      hg["b"].data_ptr(), S), "pmvae_linear_backward")
        self.part.backward(self.params, self.grads, part_acts, dctx.view_as(part_acts[-1]).contiguous(), need_dx=False)
        # decoder: Bernoulli -> conv-transpose stack -> dz
        dlogits = self.bern.backward(g_rec).view_as(dec_acts[-1]).contiguous()
        dz = self.dec.backward(self.params, self.grads, dec_acts, dlogits, need_dx=True).reshape(B, d)
        if not self._stop:
            dz = dz + dz_match
        # posterior head: (z, kl) -> par -> Linear -> conv stack
        dpar = torch.empty_like(par)
        _lib.check(_lib.lib.pmvae_tril_sample_kl_backward(par.data_ptr(), eps.data_ptr(), dz.contiguous().data_ptr(),
                                                          g_kl.data_ptr(), B, d, dpar.data_ptr(), S),
                   "pmvae_tril_sample_kl_backward")
        hw, hg = self.params["posterior_dist/linear"], self.grads["posterior_dist/linear"]
        dfeat = torch.empty_like(feat)
        _lib.check(_lib.lib.pmvae_linear_backward(feat.data_ptr(), hw["w"].data_ptr(), dpar.data_ptr(), B, self.enc_feat,
                                                  self.P, 0, dfeat.data_ptr(), hg["w"].data_ptr(), hg["b"].data_ptr(), S),
                   "pmvae_linear_backward")
        self.enc.backward(self.params, self.grads, enc_acts, dfeat.view_as(enc_acts[-1]).contiguous(), need_dx=False)
        return self.grads

    # ---- vae.py:146-226 for the MNIST config -------------------------------------------------------
    def _posterior_par(self, x):
        """raw TriLGaussian parameters of q(z | x): conv encoder + Linear head (vae.py:47-49)."""
        B = x.shape[0]
        feat = self.enc.forward(self.params, x)[-1].reshape(B, self.enc_feat)
        par = torch.empty((B, self.P), dtype=torch.float32, device=self.device)
        hw = self.params["posterior_dist/linear"]
        _lib.check(_lib.lib.pmvae_linear(_lib.PREC_F32, feat.data_ptr(), hw["w"].data_ptr(), hw["b"].data_ptr(), B,
                                         self.enc_feat, self.P, 0, par.data_ptr(), None, 0, _stream()), "pmvae_linear")
        return par

    def _context(self, x, b):
        xob = torch.cat([x * b, b], dim=-1).contiguous()
        return self.part.forward(self.params, xob)[-1].reshape(x.shape[0], -1)

    def _decode_logits(self, z):
        """decoder logits [K*B, 28, 28, 1] of latent samples z [K, B, d]."""
        K, B, d = z.shape
        return self.dec.forward(self.params, z.reshape(K * B, 1, 1, d).contiguous())[-1]

    # ---- vae.py:47-53: the sub-modules as distribution objects ------------------------------------------
    def encoder(self, x: torch.Tensor, is_training: bool = False) -> MultivariateNormalTriL:
        return MultivariateNormalTriL(self._posterior_par(_f32c(x, self.device)), self.latent_dim)

    def decoder(self, z: torch.Tensor, is_training: bool = False) -> BernoulliLogits:
        z = _f32c(z, self.device)
        return BernoulliLogits(self.dec.forward(self.params, z.reshape(z.shape[0], 1, 1, self.latent_dim))[-1])

    def partial_encoder(self, x_o_b: torch.Tensor, is_training: bool = False) -> MultivariateNormalTriL:
        """q(z | x_o) of a TriLGaussian partial posterior from the channel-concatenated [x_o, b] image (vae.py:132-134)."""
        if self.argmm is not None:
            raise NotImplementedError("the AutoregressiveGMM partial posterior has no distribution object: use "
                                      "`argmm.sample / argmm.log_prob` with `_context(x, b)`")
        x_o_b = _f32c(x_o_b, self.device)
        B = x_o_b.shape[0]
        ctx = self.part.forward(self.params, x_o_b)[-1].reshape(B, self.part_feat)
        return MultivariateNormalTriL(self._linear(ctx, "partial_posterior_dist/linear", self.part_feat), self.latent_dim)

    def impute(self, x_o: torch.Tensor, b: torch.Tensor, num_samples: int = 100, *, rng=None, key=None) -> torch.Tensor:
        """vae.py:146-169 -> [num_samples, B, 28, 28, 1]: z ~ q(z | x_o) (AR-GMM), decoder mean = sigmoid(logits)
        (tfd.Bernoulli.mean), observed pixels kept."""
        x_o, b = _f32c(x_o, self.device), _f32c(b, self.device)
        if key is None:
            if rng is None:
                raise ValueError("pass rng= or key=")
            key = prng.PRNGSequence(rng).next()        # conv nets draw no dropout keys
        x_o = x_o * b
        K, B = int(num_samples), x_o.shape[0]
        if self.argmm is not None:
            z = self.argmm.sample(self._context(x_o, b), K, key=key)
        else:
            z = self.partial_encoder(torch.cat([x_o, b], dim=-1)).sample(seed=key, sample_shape=K)
        mean = torch.sigmoid(self._decode_logits(z)).view(K, *x_o.shape)
        return torch.where(b.unsqueeze(0) != 0, x_o.unsqueeze(0), mean)

    def is_log_prob(self, x: torch.Tensor, b: torch.Tensor, num_samples: int = 100, *, rng=None, keys=None):
        """vae.py:171-226 -> (log p(x), log p(x_u | x_o)), each [B]."""
        x, b = _f32c(x, self.device), _f32c(b, self.device)
        if keys is None:
            if rng is None:
                raise ValueError("pass rng= or keys=")
            seq = prng.PRNGSequence(rng)
            keys = (seq.next(), seq.next())
        K, B, d, S = int(num_samples), x.shape[0], self.latent_dim, _stream()
        D = x[0].numel()
        par = self._posterior_par(x)
        ctx = self._context(x, b)
        # z ~ q(z | x): samples and log p(z) - log q(z | x) in one kernel
        z = torch.empty((K, B, d), dtype=torch.float32, device=self.device)
        ratio = torch.empty((K, B), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_tril_sample(par.data_ptr(), _lib.key_arg(keys[0]), B, K, B, 0, d, z.data_ptr(),
                                              ratio.data_ptr(), S), "pmvae_tril_sample")
        xk = x.reshape(1, B, D).expand(K, B, D).reshape(K * B, D).contiguous()
        ll = self.bern.log_prob(self._decode_logits(z).reshape(K * B, D), xk).view(K, B) + ratio
        # z' ~ q(z | x_o): samples, log p(z') - log q(z' | x_o), and the observed-pixel likelihood
        bk = b.reshape(1, B, D).expand(K, B, D).reshape(K * B, D).contiguous()
        if self.argmm is not None:
            z_xo = self.argmm.sample(ctx, K, key=keys[1])
            ctx_k = ctx.unsqueeze(0).expand(K, B, ctx.shape[1]).reshape(K * B, -1).contiguous()
            log_q = self.argmm.log_prob(z_xo.reshape(K * B, d), ctx_k).view(K, B)
            log_pz = torch.empty(K * B, dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib.pmvae_std_normal_log_prob(z_xo.data_ptr(), K * B, d, log_pz.data_ptr(), S),
                       "pmvae_std_normal_log_prob")
            ratio_o = log_pz.view(K, B) - log_q
        else:
            par_p = self._linear(ctx, "partial_posterior_dist/linear", self.part_feat)
            z_xo = torch.empty((K, B, d), dtype=torch.float32, device=self.device)
            ratio_o = torch.empty((K, B), dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib.pmvae_tril_sample(par_p.data_ptr(), _lib.key_arg(keys[1]), B, K, B, 0, d, z_xo.data_ptr(),
                                                  ratio_o.data_ptr(), S), "pmvae_tril_sample")
        ll_o = self.bern.log_prob(self._decode_logits(z_xo).reshape(K * B, D), xk, bk).view(K, B) + ratio_o
        ll, ll_o = ll.contiguous(), ll_o.contiguous()
        out = torch.empty((2, B), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_logmeanexp_rows(ll.data_ptr(), None, out[0].data_ptr(), B, K, S), "pmvae_logmeanexp_rows")
        _lib.check(_lib.lib.pmvae_logmeanexp_rows(ll.data_ptr(), ll_o.data_ptr(), out[1].data_ptr(), B, K, S),
                   "pmvae_logmeanexp_rows")
        return out[0], out[1]

    # ---- train_pm_vae.py:58-83 (beta = 1: the MNIST config has no beta schedule; weight_decay = 0) ------
    def train_step(self, x, b, *, rng=None, eps=None, lr_schedule=None, matching_coef: float = 1.0,
                   adam=(0.9, 0.999, 1e-8), grad_sync=None, global_rows: Optional[int] = None,
                   row_start: int = 0, sync_metrics: bool = True) -> Dict[str, float]:
        """One optimizer step.  Data parallel: every rank passes its own rows, `global_rows` = rows over all ranks (the
        cotangents are scaled by 1 / global_rows) and `grad_sync(tensors)` sums the two flat gradient arenas across ranks
        before the update (e.g. one NCCL all-reduce each).  `sync_metrics=False` skips the host read of the batch means
        (returns device tensors instead), so consecutive steps queue without a host round trip."""
        out = self(x, b, is_training=True, rng=rng, eps=eps, row_start=row_start, total_rows=global_rows)
        B = out["kl"].shape[0]
        ones = torch.full((B,), 1.0 / (global_rows or B), device=self.device)
        self.backward(-ones, ones, -matching_coef * ones)
        arenas = [(self.arena, self.grad_arena)]
        if self.argmm is not None:
            arenas.append((self.argmm.arena, self.argmm.grad_arena))
        if grad_sync is not None:
            grad_sync([g for _, g in arenas])
        lr = float(lr_schedule(self.step)) if lr_schedule else 1e-3
        for (arena, grads), m, v in zip(arenas, self.m, self.v):
            _lib.check(_lib.lib.pmvae_adamw_flat(arena.data_ptr(), grads.data_ptr(), m.data_ptr(), v.data_ptr(),
                                                 arena.numel(), self.step, lr, 0.0, adam[0], adam[1], adam[2], _stream()),
                       "pmvae_adamw_flat")
        self.step += 1
        if not sync_metrics:
            return {k: out[k].mean() for k in ("reconstruction_ll", "kl", "matching_ll")}
        rec, kl, match = (float(out[k].mean()) for k in ("reconstruction_ll", "kl", "matching_ll"))
        return {"reconstruction_ll": rec, "kl": kl, "matching_ll": match, "beta": 1.0,
                "loss": -(rec - kl) + matching_coef * (-match)}
