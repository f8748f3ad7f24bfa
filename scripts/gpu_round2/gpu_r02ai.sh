#!/bin/bash
# heads_bwd64 full-tile fast path: parity tests, bsds bench, heads kernel durations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ln_chain.py tests/test_gpu_condll_scale.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r02ai_pytest.txt
timeout 300 python bench.py --config bsds --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02ai_bsds.json 2>gpurun_out/r02ai_bsds.err
python -c "import json;d=json.load(open('gpurun_out/r02ai_bsds.json'));print('bsds',d['value'],d['ms_per_step'])"
CMD="python bench.py --config bsds --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:heads_bwd64 -s 4 -c 2 --csv --log-file gpurun_out/r02ai_heads.csv $CMD > gpurun_out/r02ai_ncu.log 2>&1
grep -E "gpu__time_duration" gpurun_out/r02ai_heads.csv | awk -F'","' '{print substr($5,1,40), $NF}'
