#!/bin/bash
# AR-GMM hidden Linears on the tensor cores (precision="bf16"): MNIST-config parity tests, MNIST bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mnist_dists.py tests/test_gpu_mnist_model.py tests/test_gpu_pm_vade.py -m gpu -q -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/r02ar_pytest.txt
timeout 400 python bench.py --config mnist --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02ar_mnist.json 2>gpurun_out/r02ar_mnist.err
python -c "import json;d=json.load(open('gpurun_out/r02ar_mnist.json'));print('mnist',d['value'],d['ms_per_step'],d['gpu_launches'])"; tail -3 gpurun_out/r02ar_mnist.err
