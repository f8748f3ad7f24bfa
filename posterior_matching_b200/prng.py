"""Host-side key handling on the JAX threefry stream (calls the C library's host
functions; mirrors `jax.random.PRNGKey/split/fold_in` and Haiku's `PRNGSequence`,
which is what `hk.next_rng_key()` advances -- vae.py:124,162,192-195)."""
from __future__ import annotations

import ctypes as C

from . import _lib


def PRNGKey(seed: int):
    seed = int(seed)
    return ((seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF)


def split(key, num: int = 2):
    out = (C.c_uint32 * (2 * num))()
    _lib.check(_lib.lib.pmvae_key_split_host(_lib.key_arg(key), num, out), "pmvae_key_split_host")
    return [(out[2 * i], out[2 * i + 1]) for i in range(num)]


def fold_in(key, data: int):
    out = (C.c_uint32 * 2)()
    _lib.check(_lib.lib.pmvae_key_fold_in_host(_lib.key_arg(key), int(data) & 0xFFFFFFFF, out), "pmvae_key_fold_in_host")
    return (out[0], out[1])


class PRNGSequence:
    """next(): key, sub = split(key); keep key, hand out sub."""

    def __init__(self, key_or_seed):
        self.key = PRNGKey(key_or_seed) if isinstance(key_or_seed, int) else (int(key_or_seed[0]), int(key_or_seed[1]))

    def next(self):
        self.key, sub = split(self.key, 2)
        return sub

    __next__ = next

    def skip(self, n: int):
        for _ in range(n):
            self.next()
        return self
