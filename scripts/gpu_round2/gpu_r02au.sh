#!/bin/bash
# TriL-partial conv PM-VAE (configs/pm_vae_mnist16.py) trainable: new tests + the conv / lookahead / mnist tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mnist16_model.py tests/test_gpu_lookahead.py tests/test_gpu_mnist_model.py tests/test_gpu_pm_vade.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/r02au_pytest.txt
