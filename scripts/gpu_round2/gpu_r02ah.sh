#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --config bsds --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 300 $CMD > gpurun_out/r02ah_plain.json 2>/dev/null && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"heads_bwd64" -s 4 -c 2 -o gpurun_out/r02ah_heads -f $CMD > gpurun_out/r02ah_ncu.log 2>&1
tail -2 gpurun_out/r02ah_ncu.log
