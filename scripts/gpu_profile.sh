#!/bin/bash
# Evidence run for profiles/: bench line, ncu launch list of the train step, full captures of the fused kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-p}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo plain failed; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:net_fwd_kernelILb0 -s 2 -c 2 -o gpurun_out/${TAG}_fwd -f $CMD > gpurun_out/${TAG}_ncu_fwd.log 2>&1
echo "fwd capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"net_bwd_kernel|net_fwd_kernel|tc_gemm_kernel" -s 60 -c 12 -o gpurun_out/${TAG}_train -f $CMD > gpurun_out/${TAG}_ncu_train.log 2>&1
echo "train capture exit $?"
