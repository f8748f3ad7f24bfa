#!/bin/bash
# MNIST-config step: launch list of the current build
mkdir -p gpurun_out
CMD="python bench.py --config mnist --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r02aq_plain.json 2>/dev/null && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 800 --csv --log-file gpurun_out/r02aq_mnist_launches.csv $CMD > gpurun_out/r02aq_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r02aq_mnist_launches.csv | head -30
