#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/ln_stress3.py > gpurun_out/r02f_stress4.txt 2>&1
cat gpurun_out/r02f_stress4.txt
