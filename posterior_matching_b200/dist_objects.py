"""Thin device-side stand-ins for the TFP objects the reference's module attributes return
(posterior_matching/models/vae.py:47-57): `model.encoder(x)` / `model.partial_encoder(x_o_b)` give a
`MultivariateNormalTriL` (distributions.py:101-113), `model.decoder(z)` a `Normal` with one shared scale
(IdentityGaussian, distributions.py:41-55), `model.prior` a standard `MultivariateNormalDiag` (vae.py:55-57).

Only the methods the reference calls are provided -- `.mean()`, `.sample(seed=, sample_shape=)`, `.log_prob()`,
`.entropy()`, `.kl_divergence(prior)` (vae.py:124-138,162-163,192-212; lookahead.py:126-133,219-222) -- each one a
libpmvae row kernel over the raw head output (no CPU path).  `seed` is a JAX PRNGKey (two uint32 words); TFP hands
it to jax.random.normal unsalted [R, SURVEY Appendix A.1], so `.sample(seed=k, sample_shape=K)` draws
eps = normal(k, [K, B, d]).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _n_samples(sample_shape) -> Tuple[int, bool]:
    """TFP's sample_shape: () -> one draw without a leading axis, n or (n,) -> [n, ...]."""
    if sample_shape is None or sample_shape == ():
        return 1, False
    if isinstance(sample_shape, int):
        return int(sample_shape), True
    shape = tuple(sample_shape)
    if len(shape) != 1:
        raise NotImplementedError("sample_shape must be () or one integer")
    return int(shape[0]), True


def _key(seed):
    if seed is None:
        raise ValueError("pass seed= (a JAX PRNGKey: two uint32 words); there is no global RNG on this path")
    return _lib.key_arg(seed)


class MultivariateNormalDiagStd:
    """tfd.MultivariateNormalDiag(zeros(d), ones(d)): the prior p(z) (vae.py:55-57)."""

    def __init__(self, d: int, device):
        self.d, self.device = int(d), device

    def mean(self) -> torch.Tensor:
        return torch.zeros(self.d, dtype=torch.float32, device=self.device)

    def log_prob(self, z: torch.Tensor) -> torch.Tensor:
        z = z.to(device=self.device, dtype=torch.float32).contiguous()
        lead = z.shape[:-1]
        n = int(math.prod(lead))
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_std_normal_log_prob(z.data_ptr(), n, self.d, out.data_ptr(), _stream()),
                   "pmvae_std_normal_log_prob")
        return out.view(lead)

    def sample(self, seed=None, sample_shape=()) -> torch.Tensor:
        n, lead = _n_samples(sample_shape)
        z = torch.empty((n, self.d), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_normal(_key(seed), n * self.d, 0, n * self.d, z.data_ptr(), _stream()), "pmvae_normal")
        return z if lead else z[0]

    def entropy(self) -> float:
        return 0.5 * self.d * (1.0 + math.log(2.0 * math.pi))


class MultivariateNormalTriL:
    """tfd.MultivariateNormalTriL(loc, FillScaleTriL(raw)) over the raw head output `parameters` [B, d + d(d+1)/2]
    of TriLGaussian (distributions.py:101-113)."""

    def __init__(self, parameters: torch.Tensor, d: int):
        self.parameters, self.d = parameters, int(d)
        self.device = parameters.device
        self.batch = parameters.shape[0]

    def mean(self) -> torch.Tensor:
        return self.parameters[:, :self.d]

    def sample(self, seed=None, sample_shape=(), *, row_start: int = 0, total_rows: Optional[int] = None) -> torch.Tensor:
        K, lead = _n_samples(sample_shape)
        B, d = self.batch, self.d
        z = torch.empty((K, B, d), dtype=torch.float32, device=self.device)
        ratio = torch.empty((K, B), dtype=torch.float32, device=self.device)
        total = B if total_rows is None else int(total_rows)
        _lib.check(_lib.lib.pmvae_tril_sample(self.parameters.data_ptr(), _key(seed), B, K, total, int(row_start), d,
                                              z.data_ptr(), ratio.data_ptr(), _stream()), "pmvae_tril_sample")
        return z if lead else z[0]

    def log_prob(self, z: torch.Tensor) -> torch.Tensor:
        """z [B, d] -> [B]; z [K, B, d] -> [K, B] (what jax.vmap(posterior.log_prob) gives, vae.py:217,220)."""
        z = z.to(device=self.device, dtype=torch.float32).contiguous()
        B, d = self.batch, self.d
        if z.shape[-2:] != (B, d):
            raise ValueError(f"expected z of shape [..., {B}, {d}]")
        zz = z.view(-1, B, d)
        out = torch.empty((zz.shape[0], B), dtype=torch.float32, device=self.device)
        for k in range(zz.shape[0]):
            _lib.check(_lib.lib.pmvae_tril_log_prob(self.parameters.data_ptr(), zz[k].data_ptr(), B, d, out[k].data_ptr(),
                                                    _stream()), "pmvae_tril_log_prob")
        return out.view(z.shape[:-1])

    def entropy(self) -> torch.Tensor:
        out = torch.empty(self.batch, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_tril_entropy(self.parameters.data_ptr(), self.batch, self.d, out.data_ptr(), _stream()),
                   "pmvae_tril_entropy")
        return out

    def kl_divergence(self, other) -> torch.Tensor:
        """KL(self || N(0, I)) (vae.py:130); the only `other` the reference passes is the prior."""
        if not isinstance(other, MultivariateNormalDiagStd):
            raise NotImplementedError("kl_divergence is provided against the standard-normal prior")
        B, d = self.batch, self.d
        eps = torch.zeros((B, d), dtype=torch.float32, device=self.device)
        z = torch.empty_like(eps)
        kl = torch.empty(B, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_tril_sample_kl(self.parameters.data_ptr(), eps.data_ptr(), B, d, z.data_ptr(),
                                                 kl.data_ptr(), _stream()), "pmvae_tril_sample_kl")
        return kl


class Normal:
    """tfd.Normal(loc, exp(log_scale)) with a scalar scale shared by every element: IdentityGaussian's output
    (distributions.py:41-55).  `log_prob` is elementwise, like the reference's (it sums afterwards, vae.py:127-128)."""

    def __init__(self, loc: torch.Tensor, log_scale: torch.Tensor):
        self.loc, self.log_scale = loc, log_scale      # log_scale: 0-d device tensor (a view of the parameter arena)
        self.device = loc.device

    def mean(self) -> torch.Tensor:
        return self.loc

    def stddev(self) -> torch.Tensor:
        return torch.exp(self.log_scale).expand_as(self.loc)

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        """x broadcasts against loc over leading axes (loc [K*B, D] or [B, D], x [B, D])."""
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        loc = self.loc.contiguous()
        D = loc.shape[-1]
        rows, rows_x = loc.numel() // D, x.numel() // D
        out = torch.empty_like(loc)
        _lib.check(_lib.lib.pmvae_normal_log_prob(x.data_ptr(), loc.data_ptr(), self.log_scale.data_ptr(), rows, rows_x, D,
                                                  out.data_ptr(), _stream()), "pmvae_normal_log_prob")
        return out

    def sample(self, seed=None, sample_shape=()) -> torch.Tensor:
        n, lead = _n_samples(sample_shape)
        eps = torch.empty((n,) + tuple(self.loc.shape), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.pmvae_normal(_key(seed), eps.numel(), 0, eps.numel(), eps.data_ptr(), _stream()), "pmvae_normal")
        out = self.loc.unsqueeze(0) + torch.exp(self.log_scale) * eps
        return out if lead else out[0]


class BernoulliLogits:
    """tfd.Independent(tfd.Bernoulli(logits)) as the convolutional decoders return it (distributions.py:20-25): `.logits`
    [B, H, W, C] and `.mean()` = sigmoid(logits) (what `impute` and the lookahead model read, vae.py:165, lookahead.py:
    132-133); the log-prob operator is `posterior_matching_b200.distributions.Bernoulli`."""

    def __init__(self, logits: torch.Tensor):
        self.logits = logits

    def mean(self) -> torch.Tensor:
        return torch.sigmoid(self.logits)
