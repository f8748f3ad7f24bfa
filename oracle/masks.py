"""Mask generators on the JAX PRNG stream: the device contract's specification.

Test infrastructure only (see oracle/__init__.py).

The reference draws masks on the host from `np.random.RandomState` streams that
are unseeded by default (posterior_matching/masking.py:11-13; each MNIST
sub-generator owns its own unseeded state, :238-246), so there is no reference
stream to be bit-exact with.  The contract below re-defines the same
generators in `jax.random` terms; the CUDA kernels (csrc/rng.cu) must match it bit
for bit, and its DISTRIBUTION is checked against the live reference's
(tests/golden/masks_reference.npz).

  bernoulli_mask   <- BernoulliMaskGenerator.call        masking.py:84-91
  mnist_mask       <- MNISTMaskGenerator                 masking.py:235-249
                      MixtureMaskGenerator.call          masking.py:39-47
                      ImageBernoulliMaskGenerator(0.5)   masking.py:94-104
                      FixedRectangleMaskGenerator x4     masking.py:143-157
                      SquareMaskGenerator(14)            masking.py:160-174
                      RectangleMaskGenerator(0.3, 1.0)   masking.py:107-140
"""
from __future__ import annotations

import numpy as np

from . import prng

MNIST_WEIGHTS = (2, 1, 1, 1, 1, 2, 2)
MNIST_MAX_ATTEMPTS = 4096
# FixedRectangleMaskGenerator(y1, x1, y2, x2) zeroes mask[y1:y2, x1:x2]
_FIXED = {1: (0, 0, 28, 14), 2: (0, 0, 14, 28), 3: (0, 14, 28, 28), 4: (14, 0, 28, 28)}


def bernoulli_mask(key, p: float, shape):
    """b = uniform(key, shape) < p as float32 (1 = observed)."""
    return prng.bernoulli(key, p, shape).astype(np.float32)


def mnist_categories(key, B: int):
    k_cat = prng.split(key, 4)[0]
    p = [np.float32(w) / np.float32(10) for w in MNIST_WEIGHTS]
    return prng.choice(k_cat, p, (B,))


def mnist_mask(key, B: int):
    """[B,28,28,1] float32 mask; see csrc/rng.cu for the same contract in prose."""
    k_cat, k_bern, k_sq, k_rect = prng.split(key, 4)
    cat = mnist_categories(key, B)
    out = np.ones((B, 28, 28, 1), dtype=np.float32)
    bern = None
    sq = None
    rect_cache = {}
    for r in range(B):
        c = int(cat[r])
        if c == 0:
            if bern is None:
                bern = prng.bernoulli(k_bern, 0.5, (B, 28, 28, 1)).astype(np.float32)
            out[r] = bern[r]
        elif c in _FIXED:
            y1, x1, y2, x2 = _FIXED[c]
            out[r, y1:y2, x1:x2] = 0
        elif c == 5:
            if sq is None:
                sq = prng.randint(k_sq, (B, 2), 0, 14)
            x, y = int(sq[r, 0]), int(sq[r, 1])
            out[r, y:y + 14, x:x + 14] = 0
        else:
            x1, x2, y1, y2 = 0, 27, 0, 27
            for t in range(MNIST_MAX_ATTEMPTS):
                if t not in rect_cache:
                    rect_cache[t] = prng.randint(prng.fold_in(k_rect, t), (B, 4), 0, 28)
                c0, c1, c2, c3 = (int(v) for v in rect_cache[t][r])
                xa, xb, ya, yb = min(c0, c1), max(c0, c1), min(c2, c3), max(c2, c3)
                if 0.3 * 784 <= (xb - xa + 1) * (yb - ya + 1) <= 1.0 * 784:
                    x1, x2, y1, y2 = xa, xb, ya, yb
                    break
            out[r, y1:y2 + 1, x1:x2 + 1] = 0
    return out
