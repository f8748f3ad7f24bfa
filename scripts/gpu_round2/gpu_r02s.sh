#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02s_bench_power.json 2> gpurun_out/r02s_bench_power.err
python -c "
import json; d=json.load(open('gpurun_out/r02s_bench_power.json')); print('power', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac_of_sustained_peak'])"
