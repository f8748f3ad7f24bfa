#!/bin/bash
# Evidence run for profiles/: GPU tests, default bench line, full ncu captures of the fused kernels (host-driven step:
# ncu --set full aborts on the graph-replayed step).  TAG = output prefix under gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-ev}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eval --no-graph"
timeout 600 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:net_fwd_kernelILb0 -s 2 -c 2 -o gpurun_out/${TAG}_fwd -f $CMD > gpurun_out/${TAG}_ncu_fwd.log 2>&1
echo "fwd capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"net_bwd_kernel|net_fwd_kernel|tn_grouped" -s 27 -c 9 -o gpurun_out/${TAG}_train -f $CMD > gpurun_out/${TAG}_ncu_train.log 2>&1
echo "train capture exit $?"
