#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, then (only if the plain runs passed) the
# ncu launch list and one full capture of the dominant kernel.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
BRC=$?
echo "bench exit $BRC"
tail -c 3000 gpurun_out/${TAG}_bench.json
if [ $BRC -eq 0 ] && [ "${NCU:-1}" = "1" ]; then
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eval --batch ${NCU_BATCH:-16384} > gpurun_out/${TAG}_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
     --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eval --batch ${NCU_BATCH:-16384} > gpurun_out/${TAG}_ncu.log 2>&1
  echo "ncu launches exit $?"
  if [ -n "${NCU_KERNEL:-}" ]; then
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL} -s 20 -c 3 \
       -o gpurun_out/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eval --batch ${NCU_BATCH:-16384} > gpurun_out/${TAG}_ncu_full.log 2>&1
    echo "ncu full exit $?"
  fi
fi
tail -5 gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_smoke.log | tail -5
