"""`PosteriorMatchingVAE` with the reference's interface, computed by libpmvae on a B200.

Reference: posterior_matching/models/vae.py:15-226 (class, from_config, __call__,
impute, is_log_prob); networks.py:138-162 and distributions.py:226-241 (registries).

What differs from the Haiku module, because there is no JAX here:
  * parameters live in one flat float32 CUDA arena owned by the model object
    (`model.params` is a dict of Haiku-named views: `encoder_net/linear_1` -> {w, b});
  * randomness is an explicit `rng` key (the key `hk.transform(...).apply(params,
    state, rng, ...)` would be given); the Haiku split chain behind
    `hk.next_rng_key()` is replayed so the same eps stream is drawn (F8);
  * gradients come from `backward()` (a VJP with per-row cotangents), not from
    `jax.value_and_grad`.
Tensors are torch CUDA tensors (device memory + stream plumbing only).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, Dict, Mapping, Optional, Tuple

import numpy as np
import torch

from . import _lib, prng
from .dist_objects import MultivariateNormalDiagStd, MultivariateNormalTriL, Normal


# ----------------------------------------------------------------------------- specs
class ResidualMLP:
    """networks.py:75-135 (spec only; the arithmetic is in csrc/)."""

    def __init__(self, residual_blocks: int = 2, hidden_units: int = 256, activation="relu",
                 activate_final: bool = True, dropout: float = 0.0, w_init=None, layer_norm: bool = False,
                 name: Optional[str] = None):
        if activation not in ("relu", None) or not activate_final or w_init is not None:
            raise NotImplementedError("the CUDA path implements relu / activate_final=True / default init")
        if dropout != 0.0:
            raise NotImplementedError("dropout > 0 is outside the hot path (SURVEY.md §2: Miniboone)")
        self.residual_blocks, self.hidden_units, self.layer_norm, self.name = residual_blocks, hidden_units, layer_norm, name


class _Dist:
    def __init__(self, event_size: int, w_init=None, b_init=None, name: Optional[str] = None):
        if w_init is not None or b_init is not None:
            raise NotImplementedError("custom initialisers are not supported")
        self.event_size, self.name = int(event_size), name


class TriLGaussian(_Dist):
    """distributions.py:87-113."""


class IdentityGaussian(_Dist):
    """distributions.py:28-55."""


_NETWORKS = {"ResidualMLP": ResidualMLP}
_DISTRIBUTIONS = {"TriLGaussian": TriLGaussian, "IdentityGaussian": IdentityGaussian}
# Bernoulli and AutoregressiveGMM exist as stand-alone device operators (distributions.py in this package); the
# module itself still needs the convolutional networks of the MNIST config to use them.
_NEXT_ROWS = {"ConvEncoder", "ConvDecoder", "Bernoulli", "AutoregressiveGMM", "DiagonalGaussian", "Independent"}


def get_network(network_type: str, network_config: Optional[Mapping[str, Any]] = None, name: Optional[str] = None):
    if network_type not in _NETWORKS:
        raise NotImplementedError(f"{network_type}: not built yet (SURVEY.md §8f next rows)")
    return _NETWORKS[network_type](**dict(network_config or {}), name=name)


def get_distribution(distribution_type: str, distribution_config: Optional[Mapping[str, Any]] = None,
                     name: Optional[str] = None):
    if distribution_type not in _DISTRIBUTIONS:
        raise NotImplementedError(f"{distribution_type}: not built yet (SURVEY.md §8f next rows)")
    return _DISTRIBUTIONS[distribution_type](**dict(distribution_config or {}), name=name)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


# ----------------------------------------------------------------------------- model
class PosteriorMatchingVAE:
    """A simple VAE augmented with a partially-observed encoder q(z|x_o) (vae.py:15-32)."""

    def __init__(self, latent_dim: int, encoder_net, decoder_net, partial_encoder_net, posterior_dist, decoder_dist,
                 partial_posterior_dist, matching_ll_stop_gradients: bool = False, name: Optional[str] = None, *,
                 precision: str = "bf16", device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("PosteriorMatchingVAE needs a CUDA device: the hot path has no CPU fallback")
        for n in (encoder_net, decoder_net, partial_encoder_net):
            if not isinstance(n, ResidualMLP):
                raise NotImplementedError("only ResidualMLP networks are on the CUDA path")
        if not (isinstance(posterior_dist, TriLGaussian) and isinstance(partial_posterior_dist, TriLGaussian)
                and isinstance(decoder_dist, IdentityGaussian)):
            raise NotImplementedError("only TriLGaussian posteriors with an IdentityGaussian decoder are on the CUDA path")
        H = encoder_net.hidden_units
        if not (decoder_net.hidden_units == H and partial_encoder_net.hidden_units == H):
            raise NotImplementedError("all three networks must share hidden_units")
        if posterior_dist.event_size != latent_dim or partial_posterior_dist.event_size != latent_dim:
            raise ValueError("posterior event_size must equal latent_dim")
        self.name = name
        self.latent_dim = int(latent_dim)
        self.num_features = decoder_dist.event_size
        self.device = torch.device("cuda" if device is None else device)
        self.precision = precision
        self._matching_ll_stop_gradients = bool(matching_ll_stop_gradients)
        self._nets = (encoder_net, decoder_net, partial_encoder_net)
        self.cfg = _lib.make_config(
            D=self.num_features, d=latent_dim, H=H, R_enc=encoder_net.residual_blocks,
            R_dec=decoder_net.residual_blocks, R_part=partial_encoder_net.residual_blocks,
            ln_enc=encoder_net.layer_norm, ln_dec=decoder_net.layer_norm, ln_part=partial_encoder_net.layer_norm,
            stop_grad=matching_ll_stop_gradients,
            precision={"fp32": _lib.PREC_F32, "bf16": _lib.PREC_BF16}[precision])
        self._cfgp = C.byref(self.cfg)
        self.n_arena = int(_lib.lib.pmvae_param_count(self._cfgp))
        if self.n_arena == 0:
            _lib.check(1, "pmvae_param_count")
        self.leaves = _lib.layout(self.cfg)
        self.arena = torch.zeros(self.n_arena, dtype=torch.float32, device=self.device)
        # gradient arena + 64 trailing floats for per-step batch statistics, so that one all-reduce carries both
        self._grad_store = torch.zeros(self.n_arena + 64, dtype=torch.float32, device=self.device)
        self.grad_arena = self._grad_store[:self.n_arena]
        self.grad_tail = self._grad_store[self.n_arena:]
        self.params = self._views(self.arena)
        self.grads = self._views(self.grad_arena)
        self._ws = None
        self._ws_key = (0, 0)
        self._params_dirty = True
        self._last = None
        self.prior = MultivariateNormalDiagStd(self.latent_dim, self.device)       # vae.py:55-57

    # ---- construction -------------------------------------------------------------
    @classmethod
    def from_config(cls, config: Mapping[str, Any], name: Optional[str] = None, **kwargs) -> "PosteriorMatchingVAE":
        """vae.py:61-118, including its fallbacks: the partial encoder defaults to the
        encoder's type/config, the partial posterior to `posterior_dist`, and both dist
        configs receive event_size = latent_dim."""
        if config["encoder_net"] == "ConvEncoder":
            # configs/pm_vae_mnist.py: convolutional networks, Bernoulli decoder, AutoregressiveGMM partial posterior --
            # composed on the host from libpmvae operators (conv_vae.py; float32 unless precision="bf16": SURVEY §8f N1)
            from .conv_vae import ConvPosteriorMatchingVAE
            prec = kwargs.pop("precision", None)
            return ConvPosteriorMatchingVAE.from_config(config, name=name, precision="bf16" if prec == "bf16" else "fp32",
                                                        **kwargs)
        encoder_net = get_network(config["encoder_net"], config.get("encoder_net_config"), name="encoder_net")
        decoder_net = get_network(config["decoder_net"], config.get("decoder_net_config"), name="decoder_net")
        partial_encoder_net = get_network(
            config.get("partial_encoder_net", config["encoder_net"]),
            config.get("partial_encoder_net_config", config.get("encoder_net_config")), name="partial_encoder_net")
        posterior_dist_config = dict(config.get("posterior_dist_config", {}) or {})
        posterior_dist_config["event_size"] = config["latent_dim"]
        partial_posterior_dist_config = dict(config.get("partial_posterior_dist_config", posterior_dist_config) or {})
        partial_posterior_dist_config["event_size"] = config["latent_dim"]
        posterior_dist = get_distribution(config["posterior_dist"], posterior_dist_config, name="posterior_dist")
        decoder_dist = get_distribution(config["decoder_dist"], config.get("decoder_dist_config"), name="decoder_dist")
        partial_posterior_dist = get_distribution(config.get("partial_posterior_dist", config["posterior_dist"]),
                                                  partial_posterior_dist_config, name="partial_posterior_dist")
        return cls(config["latent_dim"], encoder_net, decoder_net, partial_encoder_net, posterior_dist, decoder_dist,
                   partial_posterior_dist, config.get("matching_ll_stop_gradients", False), name=name, **kwargs)

    def _views(self, arena: torch.Tensor) -> Dict[str, Dict[str, torch.Tensor]]:
        out: Dict[str, Dict[str, torch.Tensor]] = {}
        for name, rows, cols, w_off, b_off in self.leaves:
            if rows == 0 and cols == 0:
                out.setdefault(name, {})["log_scale"] = arena[w_off:w_off + 1].view(())
            else:
                out.setdefault(name, {})["w"] = arena[w_off:w_off + rows * cols].view(rows, cols)
                out[name]["b"] = arena[b_off:b_off + cols]
        return out

    def init(self, seed: int = 0) -> Dict[str, Dict[str, torch.Tensor]]:
        """Haiku's defaults: w ~ TruncatedNormal(stddev = 1/sqrt(fan_in)) cut at 2 sigma,
        b = 0, log_scale = 0 (distributions.py:50-52)."""
        g = torch.Generator(device="cpu").manual_seed(int(seed))
        for name, leaf in self.params.items():
            if "w" in leaf:
                w = torch.empty(leaf["w"].shape, dtype=torch.float32)
                torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=g)
                leaf["w"].copy_(w / math.sqrt(w.shape[0]))
                leaf["b"].zero_()
            else:
                leaf["log_scale"].zero_()
        self.mark_params_changed()
        return self.params

    def load_params(self, params: Mapping[str, Mapping[str, Any]]):
        """Copies a Haiku-style `{module: {leaf: array}}` tree (e.g. TrainState.params of a
        reference checkpoint, eval_pm_vae_uci.py:79-80,113) into the arena."""
        for name, leaf in self.params.items():
            for k, dst in leaf.items():
                src = params[name][k]
                src = src if torch.is_tensor(src) else torch.as_tensor(np.asarray(src))
                dst.copy_(src.to(device=self.device, dtype=torch.float32).reshape(dst.shape))
        self.mark_params_changed()

    def mark_params_changed(self):
        self._params_dirty = True

    # ---- plumbing -------------------------------------------------------------------
    def workspace(self, B: int, K: int = 0) -> torch.Tensor:
        if self._ws is None or B > self._ws_key[0] or K > self._ws_key[1]:
            Bk, Kk = max(B, self._ws_key[0]), max(K, self._ws_key[1])
            nbytes = int(_lib.lib.pmvae_workspace_bytes(self._cfgp, Bk, Kk))
            if nbytes == 0:
                _lib.check(1, "pmvae_workspace_bytes")
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-raw.data_ptr()) % 1024          # libpmvae wants a 1024-byte aligned workspace
            self._ws_raw, self._ws = raw, raw[off:off + nbytes]
            self._ws_key = (Bk, Kk)
            self._params_dirty = True
        return self._ws

    def _prepare(self, ws: torch.Tensor):
        if self._params_dirty:
            _lib.check(_lib.lib.pmvae_prepare_params(self._cfgp, self.arena.data_ptr(), ws.data_ptr(), ws.numel(),
                                                     _stream()), "pmvae_prepare_params")
            self._params_dirty = False

    def draw_eps(self, key, B: int, *, row_start: int = 0, total_rows: Optional[int] = None) -> torch.Tensor:
        """rows [row_start, row_start+B) of jax.random.normal(key, [total_rows, d])."""
        d = self.latent_dim
        total = B if total_rows is None else int(total_rows)
        eps = torch.empty((B, d), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_normal(_lib.key_arg(key), total * d, row_start * d, B * d, eps.data_ptr(),
                                         _stream()), "pmvae_normal")
        return eps

    def call_eps_key(self, rng):
        """Key `hk.next_rng_key()` hands to `posterior.sample` in __call__ (vae.py:124):
        the encoder ResidualMLP draws one dropout key per block first (networks.py:126)."""
        return prng.PRNGSequence(rng).skip(self.cfg.R_enc).next()

    # ---- the reference's methods ------------------------------------------------------
    def __call__(self, x: torch.Tensor, b: torch.Tensor, is_training: bool = False, *, rng=None,
                 eps: Optional[torch.Tensor] = None, row_start: int = 0,
                 total_rows: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """vae.py:120-144 -> {"reconstruction_ll", "kl", "matching_ll"}, each [B].
        `is_training` only toggles dropout in the reference (networks.py:114); every
        supported config has rate 0."""
        x = _f32c(x, self.device)
        b = _f32c(b, self.device)
        B = x.shape[0]
        if x.shape != (B, self.num_features) or b.shape != x.shape:
            raise ValueError(f"expected x, b of shape [B, {self.num_features}]")
        if eps is None:
            if rng is None:
                raise ValueError("pass rng= (the key hk.transform's apply would receive) or eps=")
            eps = self.draw_eps(self.call_eps_key(rng), B, row_start=row_start, total_rows=total_rows)
        eps = _f32c(eps, self.device)
        out = torch.empty((3, B), dtype=torch.float32, device=self.device)
        ws = self.workspace(B)
        self._prepare(ws)
        _lib.check(_lib.lib.pmvae_forward(self._cfgp, self.arena.data_ptr(), x.data_ptr(), b.data_ptr(),
                                          eps.data_ptr(), B, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                          ws.data_ptr(), ws.numel(), _stream()), "pmvae_forward")
        self._last = (x, b, eps, B)
        return {"reconstruction_ll": out[0], "kl": out[1], "matching_ll": out[2]}

    def backward(self, g_rec: torch.Tensor, g_kl: torch.Tensor, g_match: torch.Tensor):
        """VJP of the last __call__: per-row cotangents -> `self.grads` (overwritten)."""
        if self._last is None:
            raise RuntimeError("backward() needs a preceding __call__")
        x, b, eps, B = self._last
        g = [_f32c(t, self.device) for t in (g_rec, g_kl, g_match)]
        ws = self.workspace(B)
        _lib.check(_lib.lib.pmvae_backward(self._cfgp, self.arena.data_ptr(), x.data_ptr(), b.data_ptr(),
                                           eps.data_ptr(), B, g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(),
                                           self.grad_arena.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
                   "pmvae_backward")
        return self.grads

    def net_apply(self, which: int, inp: torch.Tensor, msk: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Raw distribution parameters of one network + head (pmvae_net_apply): which = 0
        encoder(x) -> [B, P], 1 decoder(z) -> [B, D], 2 partial_encoder([x*b, b]) -> [B, P]."""
        inp = _f32c(inp, self.device)
        msk = _f32c(msk, self.device) if msk is not None else None
        B = inp.shape[0]
        d = self.latent_dim
        cols = self.num_features if which == 1 else d + d * (d + 1) // 2
        if out is None:
            out = torch.empty((B, cols), dtype=torch.float32, device=self.device)
        ws = self.workspace(B)
        self._prepare(ws)
        _lib.check(_lib.lib.pmvae_net_apply(self._cfgp, self.arena.data_ptr(), int(which), inp.data_ptr(),
                                            msk.data_ptr() if msk is not None else None, B, out.data_ptr(),
                                            ws.data_ptr(), ws.numel(), _stream()), "pmvae_net_apply")
        return out

    def encoder(self, x: torch.Tensor, is_training: bool = False) -> MultivariateNormalTriL:
        """`self.encoder(x)` (vae.py:47-49): q(z | x) as a distribution object; `.parameters` is the raw
        TriLGaussian head output [B, d + d(d+1)/2]."""
        return MultivariateNormalTriL(self.net_apply(0, x), self.latent_dim)

    def decoder(self, z: torch.Tensor, is_training: bool = False) -> Normal:
        """`self.decoder(z)` (vae.py:50-51): p(x | z) = Normal(loc, exp(log_scale)) (IdentityGaussian)."""
        return Normal(self.net_apply(1, z), self.params["decoder_dist"]["log_scale"])

    def partial_encoder(self, x_o_b: torch.Tensor, is_training: bool = False) -> MultivariateNormalTriL:
        """`self.partial_encoder(concat([x_o, b]))` (vae.py:52-53,132-134): takes the concatenated [B, 2D] input
        like the reference and returns q(z | x_o)."""
        D = self.num_features
        return MultivariateNormalTriL(self.net_apply(2, x_o_b[:, :D], x_o_b[:, D:]), self.latent_dim)

    def impute(self, x_o: torch.Tensor, b: torch.Tensor, num_samples: int = 100, *, rng=None, key=None,
               row_start: int = 0, total_rows: Optional[int] = None) -> torch.Tensor:
        """vae.py:146-169 -> [num_samples, B, D]: imputed values where b == 0, x_o elsewhere.  `rng` replays the Haiku
        chain of a stand-alone call (the partial encoder's dropout keys come first); `key` passes the sample key."""
        x_o = _f32c(x_o, self.device)
        b = _f32c(b, self.device)
        B = x_o.shape[0]
        if key is None:
            if rng is None:
                raise ValueError("pass rng= or key=")
            key = prng.PRNGSequence(rng).skip(self.cfg.R_part).next()
        total = B if total_rows is None else int(total_rows)
        out = torch.empty((int(num_samples), B, self.num_features), dtype=torch.float32, device=self.device)
        ws = self.workspace(B, num_samples)
        self._prepare(ws)
        _lib.check(_lib.lib.pmvae_impute(self._cfgp, self.arena.data_ptr(), x_o.data_ptr(), b.data_ptr(), B,
                                         int(num_samples), _lib.key_arg(key), total, row_start, out.data_ptr(), None,
                                         ws.data_ptr(), ws.numel(), _stream()), "pmvae_impute")
        return out

    def impute_mean(self, x_o: torch.Tensor, b: torch.Tensor, num_samples: int = 100, *, key,
                    row_start: int = 0, total_rows: Optional[int] = None) -> torch.Tensor:
        """mean over samples of `impute` (vae.py:146-169; eval_pm_vae_uci.py:88-89)."""
        x_o = _f32c(x_o, self.device)
        b = _f32c(b, self.device)
        B = x_o.shape[0]
        total = B if total_rows is None else int(total_rows)
        out = torch.empty((B, self.num_features), dtype=torch.float32, device=self.device)
        ws = self.workspace(B, num_samples)
        self._prepare(ws)
        _lib.check(_lib.lib.pmvae_impute_mean(self._cfgp, self.arena.data_ptr(), x_o.data_ptr(), b.data_ptr(), B,
                                              int(num_samples), _lib.key_arg(key), total, row_start, out.data_ptr(),
                                              ws.data_ptr(), ws.numel(), _stream()), "pmvae_impute_mean")
        return out

    def is_log_prob(self, x: torch.Tensor, b: torch.Tensor, num_samples: int = 100, *, rng=None, keys=None,
                    row_start: int = 0, total_rows: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """vae.py:171-226 -> (log p(x), log p(x_u | x_o)), each [B].  `rng` replays the
        Haiku chain of a stand-alone call (enc dropouts, partial-enc dropouts, z, z_xo);
        `keys=(key_z, key_zxo)` passes the two sample keys directly."""
        x = _f32c(x, self.device)
        b = _f32c(b, self.device)
        B = x.shape[0]
        if keys is None:
            if rng is None:
                raise ValueError("pass rng= or keys=")
            seq = prng.PRNGSequence(rng).skip(self.cfg.R_enc + self.cfg.R_part)
            keys = (seq.next(), seq.next())
        total = B if total_rows is None else int(total_rows)
        out = torch.empty((2, B), dtype=torch.float32, device=self.device)
        ws = self.workspace(B, num_samples)
        self._prepare(ws)
        _lib.check(_lib.lib.pmvae_is_log_prob(self._cfgp, self.arena.data_ptr(), x.data_ptr(), b.data_ptr(), B,
                                              int(num_samples), _lib.key_arg(keys[0]), _lib.key_arg(keys[1]), total,
                                              row_start, out[0].data_ptr(), out[1].data_ptr(), ws.data_ptr(),
                                              ws.numel(), _stream()), "pmvae_is_log_prob")
        return out[0], out[1]

    def eval_keys(self, rng):
        """Sample keys inside eval_pm_vae_uci.py's eval_fn (:82-94), replaying Haiku's
        chain: partial-enc dropouts, impute z, decoder dropouts (traced
        once under vmap), enc + partial-enc dropouts, z, z_xo."""
        seq = prng.PRNGSequence(rng).skip(self.cfg.R_part)
        k_imp = seq.next()
        seq.skip(self.cfg.R_dec + self.cfg.R_enc + self.cfg.R_part)
        return k_imp, seq.next(), seq.next()
