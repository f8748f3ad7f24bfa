#!/bin/bash
# attribution of the SAVE-mode overhead of the LayerNorm forward chain (debug bits: 16 no operand TMA stores, 64 no xhat staging, 128 no mask / rstd stores)
mkdir -p gpurun_out
for dbg in 0 16 64 128 80 208; do PMVAE_FUSED_DEBUG=$dbg timeout 120 python scripts/ln_time.py 2>&1 | tail -1; done | tee gpurun_out/r02al_ln_save_attribution.txt
