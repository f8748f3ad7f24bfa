"""CPU checks of the lookahead oracle (oracle/model_lookahead.py, oracle/prng.permutation): the properties of
jax.random.choice(replace=False) the restatement must have, the valid-mask / denominator rules of lookahead.py:158-199,
and the closed form of the objective on a hand-made case."""
import math

import numpy as np
import torch

from oracle import model as M, model_lookahead as OL, prng as P
from tests.util import conditioned_params, make_inputs, spec_of


def test_permutation_is_a_permutation_and_choice_its_prefix():
    for n in (1, 2, 8, 21, 63, 784):
        key = P.PRNGKey(100 + n)
        perm = P.permutation(key, n)
        assert sorted(perm.tolist()) == list(range(n))
        k = min(n, 16)
        assert P.choice_without_replacement(key, n, k).tolist() == perm[:k].tolist()
    # one sort round for every feature count of the reference's datasets; the order is the stable argsort of the bits
    key = P.PRNGKey(5)
    bits = P.random_bits(P.split(key, 2)[1], 8)
    assert P.permutation(key, 8).tolist() == np.argsort(bits, kind="stable").tolist()
    # different keys give different draws; the marginal of the first pick is close to uniform
    firsts = np.array([P.permutation(P.PRNGKey(s), 8)[0] for s in range(400)])
    counts = np.bincount(firsts, minlength=8)
    assert counts.min() > 25 and counts.max() < 80


def test_lookahead_objective_rules():
    spec = spec_of("gas")
    p = conditioned_params(spec)
    R, H, K, S = 2, 32, 4, 5
    lp = OL.init_params(spec, R, H)
    x, b, _ = make_inputs(spec, 6, seed=1)
    b[0] = 1.0                       # everything observed: no valid candidate -> 0
    b[1] = 0.0                       # nothing observed: every candidate valid
    key = P.PRNGKey(77)
    inds, valid, z1 = OL.model_one_step_samples(p, spec, x, b, key, K, S)
    assert len(set(inds.tolist())) == S and z1.shape == (K, 6, S, spec.d)
    assert valid[0].sum() == 0 and valid[1].sum() == S
    # a candidate is valid exactly when its feature is unobserved
    assert torch.equal(valid, 1.0 - b[:, torch.as_tensor(inds)])
    ll = OL.lookahead_lls(lp, spec, R, False, x, b, inds, valid, z1)
    assert ll[0] == 0 and torch.isfinite(ll).all()
    # closed form for row 1 from the definition
    loc, scale = OL.lookahead_encoder(lp, spec, R, False, torch.cat([x * b, b], -1))
    tot = 0.0
    for s, f in enumerate(inds.tolist()):
        for k in range(K):
            t = (z1[k, 1, s] - loc[1, f]) / scale[1, f]
            tot += float((-0.5 * t ** 2 - torch.log(scale[1, f]) - 0.5 * math.log(2 * math.pi)).sum()) / K
    assert abs(float(ll[1]) - tot / S) < 1e-9
    # same key -> same draw; the one-step samples condition on the extra feature (b_look), so they differ across s
    inds2, _, z1b = OL.model_one_step_samples(p, spec, x, b, key, K, S)
    assert inds2.tolist() == inds.tolist() and torch.equal(z1, z1b)
    g = OL.expected_info_gains(p, lp, spec, R, False, x[2], b[2])
    assert torch.isinf(g[b[2] == 1]).all() and torch.isfinite(g[b[2] == 0]).all()


def test_lookahead_oracle_matches_its_committed_fixture():
    """tests/golden/lookahead_golden.npz (make_lookahead_golden.py): the restated key order, choice without replacement
    and objective are frozen against silent drift."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lookahead_golden.npz"))
    spec = spec_of("gas")
    p = conditioned_params(spec)
    lp = OL.init_params(spec, 2, 32)
    x, b, _ = make_inputs(spec, 6, seed=1)
    loss, ll, grads, (inds, valid, z1) = OL.loss_and_grads(p, lp, spec, 2, False, x, b, P.PRNGKey(77), 4, 5)
    assert inds.tolist() == g["inds"].tolist() and np.array_equal(valid.numpy(), g["valid"])
    assert np.allclose(z1.numpy(), g["z1"], rtol=1e-12, atol=1e-12) and np.allclose(ll.numpy(), g["ll"], rtol=1e-12)
    assert abs(float(loss) - float(g["loss"])) < 1e-12
    assert np.allclose(grads[OL.HEAD]["b"].numpy(), g["g_head_b"], rtol=1e-10, atol=1e-14)
    assert P.permutation(P.PRNGKey(3), 21).tolist() == g["perm21"].tolist()
    assert P.permutation(P.PRNGKey(4), 784)[:16].tolist() == g["perm784_head"].tolist()
