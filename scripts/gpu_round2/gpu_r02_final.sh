#!/bin/bash
# round 2, final build: full GPU test suite, smoke, default bench (N = 1), reference arm, bsds bench, launch lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r02_final_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_final_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err
timeout 300 python bench.py --config bsds --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline > gpurun_out/r02_final_bench_bsds.json 2> gpurun_out/r02_final_bench_bsds.err
CMD="python bench.py --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 300 $CMD > gpurun_out/r02_final_plain.json 2>/dev/null && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02_final_power_launches.csv $CMD > gpurun_out/r02_final_ncu.log 2>&1
CMD="python bench.py --config bsds --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 300 $CMD > gpurun_out/r02_final_plain_bsds.json 2>/dev/null && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/r02_final_bsds_step_dram.csv $CMD > gpurun_out/r02_final_ncu_bsds.log 2>&1
python scripts/step_dram_summary.py gpurun_out/r02_final_bsds_step_dram.csv > gpurun_out/r02_final_bsds_step_dram_summary.txt
tail -3 gpurun_out/r02_final_pytest.txt; tail -3 gpurun_out/r02_final_smoke.txt; tail -c 300 gpurun_out/r02_final_bench.err; head -c 300 gpurun_out/r02_final_bench.json; echo; head -c 300 gpurun_out/r02_final_bench_reference.json; echo; tail -1 gpurun_out/r02_final_bsds_step_dram_summary.txt
