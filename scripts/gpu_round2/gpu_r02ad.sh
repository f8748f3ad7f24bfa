#!/bin/bash
# heads_bwd64 with 64-row blocks (vs 32), flat rec_ll backward: parity tests, bsds bench both ways, hepmass bench, bsds step DRAM list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ln_chain.py tests/test_gpu_condll_scale.py tests/test_gpu_boundary.py tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r02ad_pytest.txt
for rows in 64 32; do
PMVAE_HEADS_ROWS=$rows timeout 300 python bench.py --config bsds --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02ad_bsds_$rows.json 2>gpurun_out/r02ad_bsds.err
python -c "import json;d=json.load(open('gpurun_out/r02ad_bsds_$rows.json'));print('bsds rows=$rows',d['value'],d['ms_per_step'])"
done
timeout 300 python bench.py --config hepmass --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02ad_hepmass.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r02ad_hepmass.json'));print('hepmass',d['value'],d['ms_per_step'])"
CMD="python bench.py --config bsds --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 300 $CMD > gpurun_out/r02ad_plain.json 2>/dev/null && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/r02ad_bsds_step_dram.csv $CMD > gpurun_out/r02ad_ncu.log 2>&1
python scripts/step_dram_summary.py gpurun_out/r02ad_bsds_step_dram.csv | head -16
