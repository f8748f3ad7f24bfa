#!/bin/bash
# ncu launch list (gpu__time_duration) of a short bench run, train + eval.
set -u
mkdir -p gpurun_out
TAG=${1:-l}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-}"
timeout 600 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "launch list exit $?"
