#!/bin/bash
# heads_bwd64 with NH tiles per block: parity tests (default NH), bsds bench for NH = 1, 2, 4, step DRAM list per NH
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ln_chain.py tests/test_gpu_condll_scale.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r02ag_pytest.txt
CMD="python bench.py --config bsds --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
for nh in 1 2 4; do
PMVAE_HEADS_NH=$nh timeout 300 python bench.py --config bsds --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02ag_bsds_$nh.json 2>gpurun_out/r02ag_bsds.err
python -c "import json;d=json.load(open('gpurun_out/r02ag_bsds_$nh.json'));print('bsds nh=$nh',d['value'],d['ms_per_step'])"
PMVAE_HEADS_NH=$nh timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:heads_bwd64 -s 4 -c 4 --csv --log-file gpurun_out/r02ag_heads_$nh.csv $CMD > gpurun_out/r02ag_ncu.log 2>&1
grep -E "gpu__time_duration" gpurun_out/r02ag_heads_$nh.csv | awk -F'","' '{print $5, $NF}' | cut -c1-90
done
