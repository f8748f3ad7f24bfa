#!/bin/bash
# simplified heads_bwd64, vectorised AdamW: full GPU suite, power + bsds bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r02an_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02an_power.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r02an_power.json'));print('power',d['value'],d['ms_per_step'])"
timeout 300 python bench.py --config bsds --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02an_bsds.json 2>gpurun_out/r02an_bsds.err
python -c "import json;d=json.load(open('gpurun_out/r02an_bsds.json'));print('bsds',d['value'],d['ms_per_step'])"
CMD="python bench.py --config bsds --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"heads_bwd64|adamw" -s 4 -c 3 --csv --log-file gpurun_out/r02an_k.csv $CMD > gpurun_out/r02an_ncu.log 2>&1
grep -E "gpu__time_duration" gpurun_out/r02an_k.csv | awk -F'","' '{print substr($5,1,40), $NF}'
