"""Pins oracle/prng.py against public known answers."""
import numpy as np

from oracle import prng


def test_threefry_random123_kats():
    kats = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
            ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
            ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, want in kats:
        a, b = prng.threefry2x32(key, [ctr[0]], [ctr[1]])
        assert (int(a[0]), int(b[0])) == want


def test_split_matches_published_jax_values():
    assert prng.split(prng.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert prng.split(prng.PRNGKey(42)).tolist() == [[2465931498, 3679230171], [255383827, 267815257]]


def test_normal_matches_published_jax_values():
    assert abs(float(prng.normal(prng.PRNGKey(0), (1,))[0]) - (-0.20584226)) < 2e-7
    assert abs(float(prng.normal(prng.split(prng.PRNGKey(0))[1], (1,))[0]) - (-1.2515389)) < 2e-7


def test_random_bits_odd_and_layout():
    key = prng.PRNGKey(9)
    even = prng.random_bits(key, 10)
    a, b = prng.threefry2x32(key, np.arange(5, dtype=np.uint32), np.arange(5, 10, dtype=np.uint32))
    assert np.array_equal(even, np.concatenate([a, b]))
    odd = prng.random_bits(key, 9)  # counters padded with one 0
    a, b = prng.threefry2x32(key, np.arange(5, dtype=np.uint32), np.array([5, 6, 7, 8, 0], dtype=np.uint32))
    assert np.array_equal(odd, np.concatenate([a, b])[:9])
    assert prng.random_bits(key, 0).shape == (0,)


def test_uniform_range_and_bernoulli_half_is_msb():
    key = prng.PRNGKey(3)
    u = prng.uniform(key, (4096,))
    assert u.min() >= 0 and u.max() < 1
    bits = prng.random_bits(key, 4096)
    assert np.array_equal(prng.bernoulli(key, 0.5, (4096,)), (bits >> 31) == 0)


def test_normal_moments_and_erfinv():
    from scipy.special import erfinv
    x = prng.normal(prng.PRNGKey(7), (200000,))
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1) < 0.01 and np.isfinite(x).all()
    u = np.linspace(-0.999999, 0.999999, 20001).astype(np.float32)
    assert np.abs(prng.erfinv_f32(u) - erfinv(u.astype(np.float64))).max() < 5e-5


def test_randint_choice_ranges_and_frequencies():
    r = prng.randint(prng.PRNGKey(5), (40000,), 0, 28)
    assert r.min() == 0 and r.max() == 27
    assert np.abs(np.bincount(r, minlength=28) / r.size - 1 / 28).max() < 0.005
    p = [.2, .1, .1, .1, .1, .2, .2]
    c = prng.choice(prng.PRNGKey(6), p, (40000,))
    assert np.abs(np.bincount(c, minlength=7) / c.size - np.array(p)).max() < 0.01


def test_prng_sequence_is_split_chain():
    seq = prng.PRNGSequence(91)
    k = prng.PRNGKey(91)
    for _ in range(3):
        ks = prng.split(k)
        assert np.array_equal(seq.next(), ks[1])
        k = ks[0]


def test_normal_rows_equals_a_row_slice_of_the_full_draw():
    """`normal_rows` (used by the benchmark-scale cond-LL parity tests to draw only a chunk's eps) is a slice of
    `normal(key, (K, B, d))`, odd element counts included."""
    key = prng.PRNGKey(7)
    for K, B, d in ((3, 5, 4), (3, 5, 3), (1, 7, 1)):
        full = prng.normal(key, (K, B, d))
        for r0, nb in ((0, B), (2, 3), (B - 1, 1)):
            assert np.array_equal(prng.normal_rows(key, K, B, d, r0, nb), full[:, r0:r0 + nb])
