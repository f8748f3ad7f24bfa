#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/ln_time.py > gpurun_out/r02q_ln_time.txt 2>&1; cat gpurun_out/r02q_ln_time.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/r02q_pytest.txt; tail -4 gpurun_out/r02q_pytest.txt
timeout 600 python bench.py --config bsds --steps 10 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02q_bench_bsds.json 2> gpurun_out/r02q_bench_bsds.err
head -c 300 gpurun_out/r02q_bench_bsds.json; echo
