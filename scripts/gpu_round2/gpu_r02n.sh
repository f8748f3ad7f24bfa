#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/r02n_pytest.txt; tail -4 gpurun_out/r02n_pytest.txt
timeout 600 python bench.py --config bsds --steps 10 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02n_bench_bsds.json 2> gpurun_out/r02n_bench_bsds.err
head -c 300 gpurun_out/r02n_bench_bsds.json; echo
# power (the N = 1 workload): DRAM bytes and duration of every kernel of host-driven steps, then one full-set capture of the
# training-mode forward kernel
CMD="python bench.py --config power --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 600 $CMD > gpurun_out/r02n_plain.json 2> gpurun_out/r02n_plain.err \
 && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 200 --csv \
      --log-file gpurun_out/r02n_power_dram.csv $CMD > gpurun_out/r02n_ncu1.log 2>&1
tail -2 gpurun_out/r02n_ncu1.log
