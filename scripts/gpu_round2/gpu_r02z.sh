#!/bin/bash
# round 2 validation pass: full GPU test suite, smoke, default bench (N = 1), reference arm, bsds bench, launch lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r02z_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02z_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_bench_reference.json 2> gpurun_out/r02z_bench_reference.err
timeout 300 python bench.py --config bsds --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline > gpurun_out/r02z_bench_bsds.json 2> gpurun_out/r02z_bench_bsds.err
CMD="python bench.py --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 300 $CMD > gpurun_out/r02z_plain.json 2>/dev/null && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02z_power_launches.csv $CMD > gpurun_out/r02z_ncu.log 2>&1
tail -3 gpurun_out/r02z_pytest.txt; tail -3 gpurun_out/r02z_smoke.txt; tail -c 400 gpurun_out/r02z_bench.err; head -c 600 gpurun_out/r02z_bench.json; echo; head -c 600 gpurun_out/r02z_bench_reference.json
