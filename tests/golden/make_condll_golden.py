"""Generates tests/golden/condll_golden.npz: float64-oracle outputs of eval_fn (eval_pm_vae_uci.py:82-94) + log p(x) at
benchmark scale (B = 2048 rows, K = 512; bsds also B = 256, K = 4096), for the seeds tests/test_gpu_condll_scale.py uses.
The GPU test compares the CUDA evaluators with these vectors and re-derives a slice of each with the live oracle.

    python tests/golden/make_condll_golden.py        (CPU, ~10 minutes on 8 cores)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import model as M, prng as oprng  # noqa: E402
from tests.util import conditioned_params, make_inputs, oracle_eval_chunked, spec_of  # noqa: E402

CASES = [("gas", 2048, 512), ("power", 2048, 512), ("hepmass", 2048, 512), ("bsds", 2048, 512), ("bsds", 256, 4096)]
SEED_INPUTS, SEED_RNG = 31, 91


def main():
    out = {}
    for name, B, K in CASES:
        spec = spec_of(name)
        p = conditioned_params(spec)
        x, b, _ = make_inputs(spec, B, seed=SEED_INPUTS)
        keys = M.eval_keys(oprng.PRNGKey(SEED_RNG), spec)
        imp, ll, lpx = oracle_eval_chunked(p, spec, x, b, keys, K)
        tag = f"{name}_K{K}"
        out[tag + "_ll"] = ll.numpy()
        out[tag + "_lpx"] = lpx.numpy()
        out[tag + "_imp"] = imp.numpy().astype(np.float32)
        print(tag, float(ll.mean()), float(lpx.mean()), flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "condll_golden.npz"), **out)


if __name__ == "__main__":
    main()
