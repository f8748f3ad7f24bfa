#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-exp}
for d in 0 1 2; do PMVAE_TC_DEBUG=$d timeout 120 python scripts/gemm_bench.py; done 2>&1 | tee gpurun_out/${TAG}_gemm_bench.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'gemm',d['roofline']['achieved'],'TF', 'cond',d.get('cond_ll_eval',{}).get('value'))"
