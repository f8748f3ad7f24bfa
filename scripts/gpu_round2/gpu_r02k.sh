#!/bin/bash
mkdir -p gpurun_out
PMVAE_FUSED_DEBUG=8192 timeout 120 python scripts/ln_time.py > gpurun_out/r02k_trace.txt 2>&1
cat gpurun_out/r02k_trace.txt
