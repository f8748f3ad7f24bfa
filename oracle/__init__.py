"""CPU oracle for the PM-VAE hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (NumPy for the integer PRNG work, PyTorch
float64/float32 for the model arithmetic), what the reference computes on the
path named by BASELINE.json.  It exists to CHECK the CUDA product under
`posterior_matching_b200/`; it is never part of the product:

* only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
  `--impl reference` legs may import it;
* nothing under `posterior_matching_b200/` imports it, and the product raises
  if its CUDA library is missing.

Parity status: the reference ships NO tests, golden vectors or
fixtures for this path, and its third-party substrate (jax 0.2.26, dm-haiku
0.0.5, tfp 0.15, optax 0.1.0) is not installable here.  The oracle is
therefore pinned by public known answers instead (Random123 threefry KATs,
published `jax.random` values, the TFP `fill_triangular` doc example,
`torch.distributions` closed forms in float64, and the live reference
`masking.py` for mask distributions) -- see `tests/test_oracle_*.py` and
`tests/golden/`.  Everything that is a recollection of third-party semantics
is labelled `[R]` (unverified) in the docstrings; `[V]` means reproduced
against a public known answer.
"""
