#!/bin/bash
# last pass after the conv / lookahead generalisation: full GPU suite, smoke, MNIST bench leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r02at_pytest.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02at_smoke.txt
timeout 300 python bench.py --config mnist --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02at_mnist.json 2>gpurun_out/r02at_mnist.err
python -c "import json;d=json.load(open('gpurun_out/r02at_mnist.json'));print('mnist',d['value'],d['ms_per_step'])"; tail -2 gpurun_out/r02at_mnist.err
