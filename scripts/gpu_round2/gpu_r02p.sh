#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/ln_time.py > gpurun_out/r02p_plain.txt 2>&1 \
 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:net_bwd_ln_kernel -s 2 -c 1 -f -o gpurun_out/r02p_bwdln python scripts/ln_time.py > gpurun_out/r02p_ncu.log 2>&1
tail -3 gpurun_out/r02p_ncu.log; ls -la gpurun_out/r02p_bwdln.ncu-rep
timeout 300 python -m pytest tests/test_gpu_pm_vade.py -q -p no:cacheprovider 2>&1 | tail -15
