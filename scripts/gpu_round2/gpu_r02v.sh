#!/bin/bash
# d = 64 latent kernels after the row prefetch / paired-column rewrite: parity tests, bsds + power bench, launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_ln_chain.py tests/test_gpu_condll_scale.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02v_pytest.txt
timeout 300 python bench.py --config bsds --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02v_bsds.json 2>gpurun_out/r02v_bsds.err
python -c "import json;d=json.load(open('gpurun_out/r02v_bsds.json'));print('bsds',d['value'],d['ms_per_step'])"
CMD="python bench.py --config bsds --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 300 $CMD > gpurun_out/r02v_plain.json 2>/dev/null && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02v_bsds_launches.csv $CMD > gpurun_out/r02v_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r02v_bsds_launches.csv 2>/dev/null | head -24
