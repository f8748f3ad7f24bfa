#!/bin/bash
# one ncu --set full capture of a kernel: TAG KREGEX SKIP COUNT (bench args via BENCH_ARGS)
set -u
mkdir -p gpurun_out
TAG=${1:-n}; KREGEX=${2:-net_fwd_kernel}; SKIP=${3:-14}; COUNT=${4:-3}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-}"
timeout 600 $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s $SKIP -c $COUNT -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/${TAG}_ncu_full.log
