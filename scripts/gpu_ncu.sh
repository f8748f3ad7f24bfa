#!/bin/bash
# ncu evidence for one configuration: plain run, launch list, full capture of one kernel.
set -u
mkdir -p gpurun_out
TAG=${1:-ncu}; KREGEX=${2:-tc_gemm_kernel}; BATCH=${3:-32768}; SKIP=${4:-30}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eval --batch $BATCH"
timeout 600 $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s $SKIP -c 6 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/${TAG}_plain.log | cut -c1-600
