#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1024 2048 4096 7168; do PMVAE_FUSED_DEBUG=$dbg timeout 120 python scripts/ln_time.py >> gpurun_out/r02j_ln_time.txt 2>&1; done
cat gpurun_out/r02j_ln_time.txt
timeout 600 python bench.py --config bsds --steps 10 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02j_bench_bsds.json 2> gpurun_out/r02j_bench_bsds.err
head -c 300 gpurun_out/r02j_bench_bsds.json
