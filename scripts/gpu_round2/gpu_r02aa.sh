#!/bin/bash
# d = 64: partial-posterior solve moved into the forward.  Parity tests, bsds bench, per-kernel DRAM bytes of a bsds step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ln_chain.py tests/test_gpu_condll_scale.py tests/test_gpu_boundary.py tests/test_gpu_lookahead.py tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r02aa_pytest.txt
timeout 300 python bench.py --config bsds --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02aa_bsds.json 2>gpurun_out/r02aa_bsds.err
python -c "import json;d=json.load(open('gpurun_out/r02aa_bsds.json'));print('bsds',d['value'],d['ms_per_step'],d['roofline']['step_frac_of_sustained_peak'])"
CMD="python bench.py --config bsds --steps 2 --warmup 3 --no-all-configs --no-cpu-baseline --no-eval --no-graph"
timeout 300 $CMD > gpurun_out/r02aa_plain.json 2>/dev/null && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/r02aa_bsds_step_dram.csv $CMD > gpurun_out/r02aa_ncu.log 2>&1
tail -2 gpurun_out/r02aa_ncu.log
