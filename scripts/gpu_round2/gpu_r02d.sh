#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/ln_diag.py > gpurun_out/r02d_ln_diag_fused.txt 2>&1
PMVAE_FUSED=0 timeout 300 python scripts/ln_diag.py > gpurun_out/r02d_ln_diag_unfused.txt 2>&1
timeout 300 python -m pytest "tests/test_gpu_model.py::test_forward_matches_oracle" "tests/test_gpu_condll_scale.py" -q -p no:cacheprovider -k "bsds" 2>&1 | grep -E "^E|assert|passed|failed" | head -40 > gpurun_out/r02d_pytest.txt
cat gpurun_out/r02d_ln_diag_fused.txt gpurun_out/r02d_ln_diag_unfused.txt gpurun_out/r02d_pytest.txt
