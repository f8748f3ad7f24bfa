"""GPU parity: device threefry streams and mask generators vs the oracle (bit-exact
for bits / masks; float tolerance for erfinv-based normals)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import masks as omasks, prng as oprng

pytestmark = pytest.mark.gpu


def _lib():
    from posterior_matching_b200 import _lib as L
    return L


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _bits(key, n_total, start, count):
    L = _lib()
    out = torch.empty(max(count, 1), dtype=torch.int32, device="cuda")
    L.check(L.lib.pmvae_random_bits(L.key_arg(key), n_total, start, count, out.data_ptr(), _stream()), "bits")
    torch.cuda.synchronize()
    return out[:count].cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("n", [1, 2, 7, 8, 1001, 4096, 100003])
def test_random_bits_bit_exact_full_and_slices(n):
    key = oprng.PRNGKey(1234 + n)
    want = oprng.random_bits(key, n)
    assert np.array_equal(_bits(key, n, 0, n), want)
    if n > 4:
        s, c = n // 3, n // 2
        assert np.array_equal(_bits(key, n, s, c), want[s:s + c])
        assert np.array_equal(_bits(key, n, n - 1, 1), want[-1:])
    assert _bits(key, n, 0, 0).size == 0


def test_uniform_and_normal_match_oracle():
    L = _lib()
    key = oprng.PRNGKey(77)
    n = 65537
    u = torch.empty(n, device="cuda")
    z = torch.empty(n, device="cuda")
    L.check(L.lib.pmvae_uniform(L.key_arg(key), n, 0, n, u.data_ptr(), _stream()), "uniform")
    L.check(L.lib.pmvae_normal(L.key_arg(key), n, 0, n, z.data_ptr(), _stream()), "normal")
    assert np.array_equal(u.cpu().numpy(), oprng.uniform(key, (n,)))          # exact: integer ops + one subtract
    want = oprng.normal(key, (n,))
    got = z.cpu().numpy()
    assert np.isfinite(got).all()
    # same polynomial, different libm log1p / fma contraction: a few ulp of the result
    assert np.abs(got - want).max() < 5e-6
    # published jax value: normal(PRNGKey(0), (1,)) = -0.20584226
    one = torch.empty(1, device="cuda")
    L.check(L.lib.pmvae_normal(L.key_arg(oprng.PRNGKey(0)), 1, 0, 1, one.data_ptr(), _stream()), "normal")
    assert abs(float(one.item()) - (-0.20584226)) < 1e-6


@pytest.mark.parametrize("B,D,p", [(1, 1, 0.5), (64, 8, 0.5), (513, 21, 0.5), (300, 63, 0.3), (0, 8, 0.5)])
def test_bernoulli_mask_bit_exact(B, D, p):
    from posterior_matching_b200.masking import BernoulliMaskGenerator
    key = oprng.PRNGKey(5)
    gen = BernoulliMaskGenerator(p=p, seed=0)
    got = gen((B, D), key=tuple(int(k) for k in key)).cpu().numpy()
    want = omasks.bernoulli_mask(key, p, (B, D))
    assert got.dtype == np.float32 and np.array_equal(got, want)
    if B >= 4:  # a rank's row slice of the global draw equals the slice of the full draw
        part = gen((B // 2, D), key=tuple(int(k) for k in key), row_start=B // 4, total_rows=B).cpu().numpy()
        assert np.array_equal(part, want[B // 4:B // 4 + B // 2])


def test_bernoulli_mask_generator_stream_is_seeded_and_advances():
    from posterior_matching_b200.masking import get_mask_generator
    a = get_mask_generator("BernoulliMaskGenerator", seed=3)
    b = get_mask_generator("BernoulliMaskGenerator", seed=3)
    m1, m2 = a((128, 8)), a((128, 8))
    assert torch.equal(m1, b((128, 8))) and not torch.equal(m1, m2)
    assert abs(float(m1.mean()) - 0.5) < 0.1
    with pytest.raises(KeyError):
        get_mask_generator("MixtureMaskGenerator")      # (masking.py:328-335 names more generators than the configs use)


def test_mnist_mask_bit_exact_and_reference_distribution(golden_dir):
    import os
    from posterior_matching_b200.masking import MNISTMaskGenerator
    key = oprng.PRNGKey(21)
    gen = MNISTMaskGenerator(seed=0)
    B = 96
    got = gen((B, 28, 28, 1), key=tuple(int(k) for k in key)).cpu().numpy()
    want = omasks.mnist_mask(key, B)
    assert got.shape == (B, 28, 28, 1) and np.array_equal(got, want)
    part = gen((32, 28, 28, 1), key=tuple(int(k) for k in key), row_start=40, total_rows=B).cpu().numpy()
    assert np.array_equal(part, want[40:72])
    # distribution vs the live reference generator (fixture from tests/golden/make_golden.py)
    g = np.load(os.path.join(golden_dir, "masks_reference.npz"))
    n = 20000
    m = gen((n, 28, 28, 1), key=(1, 2)).cpu().numpy()[..., 0]
    cats = omasks.mnist_categories(np.array([1, 2], dtype=np.uint32), n)
    freq = np.bincount(cats, minlength=7) / n
    assert np.abs(freq - g["mnist_cat_freq"]).max() < 0.015
    assert abs(m[cats == 0].mean() - float(g["mnist_bern_mean"])) < 0.01
    area = (1 - m[cats == 6]).sum((1, 2))
    assert area.min() >= 236 and area.max() <= 784 and area.min() <= float(g["mnist_rect_area_min"]) + 10
    sq = m[cats == 5]
    assert np.all((1 - sq).sum((1, 2)) == 196)
    for c, (y1, x1, y2, x2) in {1: (0, 0, 28, 14), 2: (0, 0, 14, 28), 3: (0, 14, 28, 28), 4: (14, 0, 28, 28)}.items():
        ref = np.ones((28, 28), dtype=np.float32)
        ref[y1:y2, x1:x2] = 0
        assert np.all(m[cats == c] == ref)


def test_uniform_mask_generator_counts_and_subsets():
    """masking.py:50-81 on the device: per row a count q in [int(d lo), int(d lo) + int(d hi)) of observed features, chosen
    without replacement; seeded, advancing, shardable by rows."""
    from posterior_matching_b200 import get_mask_generator
    gen = get_mask_generator("UniformMaskGenerator", bounds=(0.0, 0.2), seed=5)
    m = gen((512, 16, 16, 1))
    assert m.shape == (512, 16, 16, 1) and set(torch.unique(m).tolist()) <= {0.0, 1.0}
    q = m.view(512, -1).sum(1)
    assert int(q.min()) >= 0 and int(q.max()) <= int(256 * 0.2) - 1        # l + choice(h): at most h - 1 above l
    assert q.float().std() > 5                                            # the count itself is spread over its range
    # every feature is picked about equally often
    freq = m.view(512, -1).mean(0)
    assert float(freq.max()) < 0.25 and float(freq.min()) > 0.01
    # same seed -> same stream; the stream advances; a row shard equals the rows of the whole batch
    gen2 = get_mask_generator("UniformMaskGenerator", bounds=(0.0, 0.2), seed=5)
    assert torch.equal(gen2((512, 16, 16, 1)), m)
    assert not torch.equal(gen2((512, 16, 16, 1)), m)
    key = (1, 2)
    whole = gen((64, 21), key=key)
    part = gen((16, 21), key=key, row_start=32, total_rows=64)
    assert torch.equal(whole[32:48], part)
    # no bounds: q in [0, d)
    free = get_mask_generator("UniformMaskGenerator", seed=1)((256, 30))
    assert int(free.sum(1).max()) <= 29
