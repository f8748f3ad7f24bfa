#!/bin/bash
# last pass: full GPU suite (incl. UniformMaskGenerator, mnist16 model), smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/r02av_pytest.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02av_smoke.txt
