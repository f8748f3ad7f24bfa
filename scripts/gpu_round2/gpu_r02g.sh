#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/ln_stress2.py bsds 150 > gpurun_out/r02g_stress.txt 2>&1
cat gpurun_out/r02g_stress.txt
if grep -q FLAKY gpurun_out/r02g_stress.txt; then echo "still flaky"; exit 0; fi
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02g_pytest.txt
tail -12 gpurun_out/r02g_pytest.txt
timeout 600 python bench.py --config bsds --steps 10 --warmup 3 --no-all-configs --no-cpu-baseline > gpurun_out/r02g_bench_bsds.json 2> gpurun_out/r02g_bench_bsds.err \
 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02g_bsds_launches.csv \
      python bench.py --config bsds --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph > gpurun_out/r02g_ncu.log 2>&1
tail -c 300 gpurun_out/r02g_bench_bsds.err; head -c 600 gpurun_out/r02g_bench_bsds.json; tail -3 gpurun_out/r02g_ncu.log
