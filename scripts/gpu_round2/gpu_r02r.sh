#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/r02r_pytest.txt; tail -4 gpurun_out/r02r_pytest.txt
timeout 300 python scripts/ln_stress.py gas 20 > gpurun_out/r02r_stress_gas.txt 2>&1; cat gpurun_out/r02r_stress_gas.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02r_bench_power.json 2> gpurun_out/r02r_bench_power.err
python -c "
import json; d=json.load(open('gpurun_out/r02r_bench_power.json')); print('power', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['step_frac_of_sustained_peak'])"
