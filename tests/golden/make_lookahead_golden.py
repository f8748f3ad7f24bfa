"""Regression fixture for oracle/model_lookahead.py and oracle/prng.permutation (run from the repo root:
`python tests/golden/make_lookahead_golden.py`).  Not a pin against JAX -- none is installable here -- but it freezes the
restated key order / choice-without-replacement / objective so that later edits of the oracle cannot drift silently."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import model_lookahead as OL, prng as P  # noqa: E402
from tests.util import conditioned_params, make_inputs, spec_of  # noqa: E402


def main():
    spec = spec_of("gas")
    p = conditioned_params(spec)
    R, H, K, S = 2, 32, 4, 5
    lp = OL.init_params(spec, R, H)
    x, b, _ = make_inputs(spec, 6, seed=1)
    key = P.PRNGKey(77)
    loss, ll, grads, (inds, valid, z1) = OL.loss_and_grads(p, lp, spec, R, False, x, b, key, K, S)
    out = {"inds": np.asarray(inds), "valid": valid.numpy(), "z1": z1.numpy(), "ll": ll.numpy(), "loss": float(loss),
           "g_head_b": grads[OL.HEAD]["b"].numpy(), "perm21": P.permutation(P.PRNGKey(3), 21),
           "perm784_head": P.permutation(P.PRNGKey(4), 784)[:16]}
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lookahead_golden.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
