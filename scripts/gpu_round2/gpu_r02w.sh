#!/bin/bash
# heads_bwd64 pair fix: parity tests, then one `ncu --set full` capture of the d = 64 latent kernels.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_ln_chain.py tests/test_gpu_condll_scale.py tests/test_gpu_pm_vade.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r02w_pytest.txt
CMD="python bench.py --config bsds --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 300 $CMD > gpurun_out/r02w_plain.json 2>/dev/null && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"heads_bwd64|solve64_bwd|match_fwd64" -s 6 -c 3 -o gpurun_out/r02w_latent64 -f $CMD > gpurun_out/r02w_ncu.log 2>&1
tail -2 gpurun_out/r02w_ncu.log
