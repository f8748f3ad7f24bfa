#!/bin/bash
# pytest (GPU) + bench for the current build; PMVAE_* env knobs pass through.
set -u
mkdir -p gpurun_out
TAG=${1:-f}
timeout 900 python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
tail -5 gpurun_out/${TAG}_bench.err
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'gemm',d['roofline']['achieved'],'TF', 'cond',d.get('cond_ll_eval',{}).get('value'))"
