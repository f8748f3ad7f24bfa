"""NumPy restatement of the JAX PRNG stream the reference draws from.

Test infrastructure only (see oracle/__init__.py).

Reference call sites (all through `hk.next_rng_key()`):
  posterior_matching/models/vae.py:124     z = posterior.sample(seed=...)
  posterior_matching/models/vae.py:162     impute: z ~ q(z|x_o) [K,B,d]
  posterior_matching/models/vae.py:192-195 is_log_prob: z, z_xo [K,B,d]
  posterior_matching/models/networks.py:126 dropout key drawn even at rate 0

Third-party algorithm restated: jax==0.2.26 `jax.random` with the default
`threefry2x32` implementation (requirements.txt:8), dm-haiku==0.0.5
`PRNGSequence` (requirements.txt:5).  SURVEY.md Appendix A.1.

[V] = reproduced against a public known answer in tests/test_oracle_prng.py.
[R] = recollection, unverified.
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds.  [V] Random123 KATs.

    key: (k0, k1) uint32 scalars; x0, x1: uint32 arrays (same shape).
    """
    k0, k1 = U32(key[0]), U32(key[1])
    x0 = np.asarray(x0, dtype=U32).copy()
    x1 = np.asarray(x1, dtype=U32).copy()
    ks = (k0, k1, U32(k0 ^ k1 ^ U32(0x1BD11BDA)))
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 += x1
                x1 = _rotl(x1, r)
                x1 ^= x0
            x0 += ks[(g + 1) % 3]
            x1 += ks[(g + 2) % 3] + U32(g + 1)
    return x0, x1


def PRNGKey(seed: int):
    """[V] `jax.random.PRNGKey(seed)` = [seed >> 32, seed & 0xFFFFFFFF]."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def random_bits(key, n: int):
    """[V] 32-bit `random_bits` for a flat count n (row-major over the shape).

    counters iota(n) padded to even, split in halves (c0, c1); one threefry
    call per pair; output = concat(first words, second words)[:n].
    """
    n = int(n)
    if n == 0:
        return np.zeros((0,), dtype=U32)
    m = n + (n & 1)
    h = m // 2
    c = np.arange(m, dtype=np.uint64).astype(U32)
    if n & 1:
        c[-1] = 0
    a, b = threefry2x32(key, c[:h], c[h:])
    return np.concatenate([a, b])[:n]


def random_bits_at(key, n: int, idx):
    """Elements `idx` (flat indices) of `random_bits(key, n)` without forming the whole draw: element i < h is the
    first word of threefry(key; i, i + h), element i >= h the second word of threefry(key; i - h, i), h = ceil(n / 2)
    (the counter appended for odd n is 0)."""
    n = int(n)
    idx = np.asarray(idx, dtype=np.uint64)
    h = (n + (n & 1)) // 2
    first = idx < h
    c0 = np.where(first, idx, idx - h)
    c1 = c0 + np.uint64(h)
    if n & 1:
        c1 = np.where(c1 == n, 0, c1)          # the padding counter
    a, b = threefry2x32(key, c0.astype(U32), c1.astype(U32))
    return np.where(first, a, b)


def normal_rows(key, K: int, B: int, d: int, row_start: int, rows: int):
    """Rows [row_start, row_start + rows) of `normal(key, (K, B, d))` -> [K, rows, d] float32."""
    k = np.arange(K, dtype=np.uint64)[:, None, None]
    r = (np.uint64(row_start) + np.arange(rows, dtype=np.uint64))[None, :, None]
    j = np.arange(d, dtype=np.uint64)[None, None, :]
    idx = ((k * np.uint64(B) + r) * np.uint64(d) + j).reshape(-1)
    bits = random_bits_at(key, K * B * d, idx)
    f = ((bits >> U32(9)) | U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = np.maximum(lo, f * (np.float32(1.0) - lo) + lo).astype(np.float32)
    return (np.float32(np.sqrt(2.0)) * erfinv_f32(u)).astype(np.float32).reshape(K, rows, d)


def split(key, num: int = 2):
    """[V] `jax.random.split(key, num)` -> [num, 2] uint32."""
    return random_bits(key, 2 * num).reshape(num, 2)


def fold_in(key, data: int):
    """[R] `jax.random.fold_in(key, data)` = threefry2x32(key; 0, data)."""
    a, b = threefry2x32(key, np.array([0], dtype=U32), np.array([data], dtype=U32))
    return np.array([a[0], b[0]], dtype=U32)


def uniform(key, shape, minval=0.0, maxval=1.0):
    """[V] float32 uniform on [minval, maxval)."""
    n = int(np.prod(shape)) if len(shape) else 1
    bits = random_bits(key, n)
    f = ((bits >> U32(9)) | U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.float32(minval)
    hi = np.float32(maxval)
    out = np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)
    return out.reshape(shape)


# Giles' single-precision erfinv polynomial, as used by XLA's ErfInv for F32
# ([R] w = -log1p(-x*x); older XLA builds used -log((1-x)(1+x))).
_ERFINV_LT5 = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
               0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941)
_ERFINV_GE5 = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
               0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682)


def erfinv_f32(x):
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (-np.log1p((-x * x).astype(np.float32))).astype(np.float32)
        lt = w < np.float32(5.0)
        w1 = (w - np.float32(2.5)).astype(np.float32)
        w2 = (np.sqrt(w) - np.float32(3.0)).astype(np.float32)
        p1 = np.full_like(x, np.float32(_ERFINV_LT5[0]))
        for c in _ERFINV_LT5[1:]:
            p1 = (np.float32(c) + p1 * w1).astype(np.float32)
        p2 = np.full_like(x, np.float32(_ERFINV_GE5[0]))
        for c in _ERFINV_GE5[1:]:
            p2 = (np.float32(c) + p2 * w2).astype(np.float32)
        p = np.where(lt, p1, p2)
        out = (p * x).astype(np.float32)
    return np.where(np.abs(x) == 1.0, np.copysign(np.float32(np.inf), x), out).astype(np.float32)


def normal(key, shape):
    """[V] `jax.random.normal(key, shape, float32)`; last digits depend on erfinv."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = uniform(key, shape, lo, 1.0)
    return (np.float32(np.sqrt(2.0)) * erfinv_f32(u)).astype(np.float32)


def bernoulli(key, p, shape):
    """[R] `jax.random.bernoulli` = uniform(key, shape) < p  (bool)."""
    return uniform(key, shape) < np.float32(p)


def randint(key, shape, minval: int, maxval: int):
    """[R] `jax.random.randint` for int32 / 32-bit draws."""
    n = int(np.prod(shape)) if len(shape) else 1
    k1, k2 = split(key)
    hi_b = random_bits(k1, n).astype(np.uint64)
    lo_b = random_bits(k2, n).astype(np.uint64)
    span = np.uint64(maxval - minval)
    mult = np.uint64((1 << 16)) % span
    mult = (mult * mult) % span
    # all uint32 arithmetic in JAX: emulate the 32-bit wrap of mul/add
    off = (((hi_b % span) * mult) & np.uint64(0xFFFFFFFF))
    off = (off + (lo_b % span)) & np.uint64(0xFFFFFFFF)
    off = off % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32).reshape(shape)


def choice_cumsum(p):
    """Sequential float32 cumsum of p (the table both oracle and device use)."""
    acc = np.float32(0.0)
    out = []
    for v in p:
        acc = np.float32(acc + np.float32(v))
        out.append(acc)
    return np.array(out, dtype=np.float32)


def choice(key, p, shape):
    """[R] `jax.random.choice(key, len(p), shape, p=p)` (with replacement)."""
    cum = choice_cumsum(p)
    u = uniform(key, shape)
    r = (cum[-1] * (np.float32(1.0) - u)).astype(np.float32)
    return np.searchsorted(cum, r, side="left").astype(np.int32)


class PRNGSequence:
    """[R] Haiku `PRNGSequence`: next() -> (key, sub) = split(key); return sub."""

    def __init__(self, key_or_seed):
        if isinstance(key_or_seed, (int, np.integer)):
            key_or_seed = PRNGKey(int(key_or_seed))
        self.key = np.asarray(key_or_seed, dtype=U32)

    def next(self):
        ks = split(self.key, 2)
        self.key = ks[0]
        return ks[1]

    __next__ = next

    def take(self, n):
        return [self.next() for _ in range(n)]


def permutation(key, n: int):
    """[R: jax 0.2.26 `jax.random.permutation(key, n)` -> `_shuffle(key, arange(n), 0)`] ceil(3 ln n / ln(2^32 - 1))
    rounds; each round: key, subkey = split(key); stable sort of the current order by 32 random bits per element
    drawn from subkey (`lax.sort_key_val`, stable by default)."""
    x = np.arange(int(n))
    rounds = int(np.ceil(3 * np.log(max(1, int(n))) / np.log(np.iinfo(np.uint32).max)))
    key = np.asarray(key, dtype=U32)
    for _ in range(rounds):
        ks = split(key, 2)
        key, sub = ks[0], ks[1]
        x = x[np.argsort(random_bits(sub, int(n)), kind="stable")]
    return x


def choice_without_replacement(key, n: int, k: int):
    """[R: jax 0.2.26 `jax.random.choice(key, n, (k,), replace=False)` with p=None] = permutation(key, n)[:k]."""
    return permutation(key, n)[: int(k)]
