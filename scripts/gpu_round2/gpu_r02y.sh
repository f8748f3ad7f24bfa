#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lookahead.py -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/r02y_pytest.txt
