"""Mask generators on the device (reference: posterior_matching/masking.py).

Same registry and call shape as the reference (`get_mask_generator(name, **kw)`,
`gen(shape) -> float32 mask`, 1 = observed) but the bits come from the JAX threefry
stream on the GPU instead of a host MT19937 (`masking.py:13`): a generator owns a key,
and every call folds a call counter into it, so the stream is reproducible for a seed
and shardable by rows (`row_start`, `total_rows`).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib, prng


def _stream():
    return torch.cuda.current_stream().cuda_stream


class MaskGenerator:
    def __init__(self, seed: Optional[int] = None, dtype=torch.float32, device=None):
        if seed is None:
            seed = int.from_bytes(__import__("os").urandom(4), "little")
        self._key = prng.PRNGKey(seed)
        self._calls = 0
        self._dtype = dtype
        self._device = torch.device("cuda" if device is None else device)

    def next_key(self):
        k = prng.fold_in(self._key, self._calls)
        self._calls += 1
        return k

    def __call__(self, shape: Sequence[int], *, key=None, row_start: int = 0, total_rows: Optional[int] = None):
        key = self.next_key() if key is None else key
        out = self.call(tuple(int(s) for s in shape), key, int(row_start), total_rows)
        return out if self._dtype == torch.float32 else out.to(self._dtype)

    def call(self, shape, key, row_start, total_rows):
        raise NotImplementedError


class BernoulliMaskGenerator(MaskGenerator):
    """masking.py:84-91: iid Bernoulli(p) per feature."""

    def __init__(self, p: float = 0.5, **kwargs):
        super().__init__(**kwargs)
        self.p = p

    def call(self, shape, key, row_start, total_rows):
        rows = shape[0]
        D = 1
        for s in shape[1:]:
            D *= s
        total = rows if total_rows is None else int(total_rows)
        out = torch.empty(shape, dtype=torch.float32, device=self._device)
        _lib.check(_lib.lib.pmvae_mask_bernoulli(_lib.key_arg(key), float(self.p), total, row_start, rows, D,
                                                 out.data_ptr(), _stream()), "pmvae_mask_bernoulli")
        return out


class MNISTMaskGenerator(MaskGenerator):
    """masking.py:235-249: per-row mixture (weights 2:1:1:1:1:2:2) of ImageBernoulli(0.5),
    four half-image rectangles, a 14x14 square and a random rectangle of 30-100% area."""

    def __init__(self, dim: int = 28, **kwargs):
        super().__init__(**kwargs)
        if dim != 28:
            raise ValueError("the device MNIST mask kernel is specialised for 28x28 images")

    def call(self, shape, key, row_start, total_rows):
        if len(shape) != 4 or tuple(shape[1:3]) != (28, 28):
            raise AssertionError(f"expected shape [batch, 28, 28, channels], got {shape}")
        rows = shape[0]
        total = rows if total_rows is None else int(total_rows)
        out = torch.empty((rows, 28, 28, 1), dtype=torch.float32, device=self._device)
        _lib.check(_lib.lib.pmvae_mask_mnist(_lib.key_arg(key), total, row_start, rows, out.data_ptr(), _stream()),
                   "pmvae_mask_mnist")
        return out


class UniformMaskGenerator(MaskGenerator):
    """masking.py:50-81: per row, a number q of observed features drawn uniformly (from [int(d lo), int(d lo) + int(d hi))
    with `bounds=(lo, hi)`, else from [0, d)), then q features chosen without replacement -- the generator of
    configs/pm_vae_mnist16.py / lookahead_mnist16.py.  On the threefry stream: q from uniform(split(key)[0]) by global
    row, the subset = the q smallest of d random 32-bit words per row (split(key)[1], global element index); the sort is
    a library call (rows of a few hundred elements, outside the hot path)."""

    def __init__(self, bounds=None, **kwargs):
        super().__init__(**kwargs)
        self._bounds = None if bounds is None else (float(bounds[0]), float(bounds[1]))

    def call(self, shape, key, row_start, total_rows):
        rows = shape[0]
        d = 1
        for s in shape[1:]:
            d *= s
        total = rows if total_rows is None else int(total_rows)
        kq, kb = prng.split(key, 2)
        u = torch.empty(rows, dtype=torch.float32, device=self._device)
        _lib.check(_lib.lib.pmvae_uniform(_lib.key_arg(kq), total, row_start, rows, u.data_ptr(), _stream()), "pmvae_uniform")
        if self._bounds is None:
            lo, span = 0, d
        else:
            lo, span = int(d * self._bounds[0]), int(d * self._bounds[1])
        q = lo + torch.clamp((u * span).floor().to(torch.int64), max=max(span - 1, 0))
        bits = torch.empty((rows, d), dtype=torch.int32, device=self._device)
        _lib.check(_lib.lib.pmvae_random_bits(_lib.key_arg(kb), total * d, row_start * d, rows * d, bits.data_ptr(), _stream()),
                   "pmvae_random_bits")
        order = torch.sort(bits.to(torch.int64) & 0xFFFFFFFF, dim=1, stable=True).indices
        picked = (torch.arange(d, device=self._device).unsqueeze(0) < q.unsqueeze(1)).to(torch.float32)
        out = torch.zeros((rows, d), dtype=torch.float32, device=self._device)
        out.scatter_(1, order, picked)
        return out.view(shape)


_GENERATORS = {
    "BernoulliMaskGenerator": BernoulliMaskGenerator,
    "MNISTMaskGenerator": MNISTMaskGenerator,
    "UniformMaskGenerator": UniformMaskGenerator,
}


def get_mask_generator(mask_generator_name: str, **kwargs) -> MaskGenerator:
    """masking.py:328-335 (the generators the PM-VAE / lookahead configs name; the others raise)."""
    if mask_generator_name not in _GENERATORS:
        raise KeyError(f"{mask_generator_name} is outside the PM-VAE hot path (SURVEY.md §2)")
    return _GENERATORS[mask_generator_name](**kwargs)
