#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/ln_time.py > gpurun_out/r02t_plain.txt 2>&1 \
 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:net_fwd_kernel -s 3 -c 1 -f -o gpurun_out/r02t_fwdln python scripts/ln_time.py > gpurun_out/r02t_ncu.log 2>&1
cat gpurun_out/r02t_plain.txt; tail -2 gpurun_out/r02t_ncu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02t_bench_power.json 2> gpurun_out/r02t_bench_power.err
python -c "
import json; d=json.load(open('gpurun_out/r02t_bench_power.json')); print('power', d['value'], d['ms_per_step'])"
