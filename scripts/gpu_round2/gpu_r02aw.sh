#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_prng.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r02aw_pytest.txt
