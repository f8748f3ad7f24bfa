"""tests/golden/condll_golden.npz (benchmark-scale cond-LL / imputation vectors of the float64 oracle, made by
tests/golden/make_condll_golden.py) is what oracle/ computes today: a slice of every case is re-derived live (CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import model as M, prng as oprng
from tests.util import conditioned_params, make_inputs, spec_of

CASES = [("gas", 2048, 512), ("power", 2048, 512), ("hepmass", 2048, 512), ("bsds", 2048, 512), ("bsds", 256, 4096)]
SEED_INPUTS, SEED_RNG = 31, 91       # tests/golden/make_condll_golden.py


@pytest.mark.parametrize("name,B,K", CASES)
def test_golden_slice_is_the_live_oracle(name, B, K, golden_dir):
    """Rows [r0, r0 + n) of the fixture recomputed now."""
    g = np.load(os.path.join(golden_dir, "condll_golden.npz"))
    tag = f"{name}_K{K}"
    spec = spec_of(name)
    p = conditioned_params(spec)
    x, b, _ = make_inputs(spec, B, seed=SEED_INPUTS)
    keys = M.eval_keys(oprng.PRNGKey(SEED_RNG), spec)
    r0, n = 100, (8 if K > 512 else 16)
    k_imp, k_z, k_zxo = keys
    e = [torch.tensor(oprng.normal_rows(k, K, B, spec.d, r0, n), dtype=torch.float64) for k in (k_imp, k_z, k_zxo)]
    with torch.no_grad():
        imp, ll = M.eval_fn(p, spec, x[r0:r0 + n], b[r0:r0 + n], *e)
    assert np.allclose(ll.numpy(), g[tag + "_ll"][r0:r0 + n], rtol=1e-9, atol=1e-9)
    assert np.allclose(imp.numpy(), g[tag + "_imp"][r0:r0 + n], rtol=1e-5, atol=1e-6)
