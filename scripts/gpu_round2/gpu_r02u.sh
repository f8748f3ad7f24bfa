#!/bin/bash
mkdir -p gpurun_out
PMVAE_FUSED_DEBUG=8192 timeout 120 python scripts/ln_time.py > gpurun_out/r02u_trace.txt 2>&1
grep -A9 "pass 1 of epilogue" gpurun_out/r02u_trace.txt | head -12; grep -A14 "^net_bwd_ln trace" gpurun_out/r02u_trace.txt | head -16
CMD="python bench.py --config bsds --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 300 $CMD > gpurun_out/r02u_plain.json 2>/dev/null && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02u_bsds_launches.csv $CMD > gpurun_out/r02u_ncu.log 2>&1
tail -1 gpurun_out/r02u_ncu.log
