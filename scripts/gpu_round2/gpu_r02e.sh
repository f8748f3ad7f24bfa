#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/ln_stress.py bsds 30 > gpurun_out/r02e_stress_bsds.txt 2>&1
timeout 200 python scripts/ln_stress.py gas 20 > gpurun_out/r02e_stress_gas.txt 2>&1
cat gpurun_out/r02e_stress_bsds.txt gpurun_out/r02e_stress_gas.txt
