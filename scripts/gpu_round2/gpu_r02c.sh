#!/bin/bash
# round 2, call 3: fused LayerNorm training chains (bsds) -- tests first (bounded), then the bsds bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ln_chain.py tests/test_gpu_nets.py -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r02c_pytest_ln.txt
tail -15 gpurun_out/r02c_pytest_ln.txt
if grep -q "failed\|error" gpurun_out/r02c_pytest_ln.txt; then echo "LN tests failed; skipping the rest"; exit 0; fi
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r02c_pytest.txt
tail -8 gpurun_out/r02c_pytest.txt
timeout 600 python bench.py --config bsds --steps 10 --warmup 3 --no-all-configs --no-cpu-baseline > gpurun_out/r02c_bench_bsds.json 2> gpurun_out/r02c_bench_bsds.err
tail -c 400 gpurun_out/r02c_bench_bsds.err; head -c 900 gpurun_out/r02c_bench_bsds.json
