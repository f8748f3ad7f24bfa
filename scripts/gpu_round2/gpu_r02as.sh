#!/bin/bash
# Lookahead over the convolutional mnist16 model; the MNIST-config tests still pass with the generalised conv class
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_mnist_model.py tests/test_gpu_pm_vade.py tests/test_gpu_boundary.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/r02as_pytest.txt
