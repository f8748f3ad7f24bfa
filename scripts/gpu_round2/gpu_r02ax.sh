#!/bin/bash
# last GPU seconds of the round: the default bench line of the final tree (CPU baseline skipped to fit the budget)
mkdir -p gpurun_out
timeout 75 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r02_last_bench.json 2> gpurun_out/r02_last_bench.err
echo "rc=$?"; head -c 400 gpurun_out/r02_last_bench.json; echo; tail -c 300 gpurun_out/r02_last_bench.err
