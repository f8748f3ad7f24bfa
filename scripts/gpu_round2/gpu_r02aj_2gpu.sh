#!/bin/bash
# 2 GPUs, final build: 2-rank tests, N = 2 bench lines for hepmass (the N > 1 workload) and bsds, reference arm under torchrun
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/r02aj_pytest_dist.txt
tail -3 gpurun_out/r02aj_pytest_dist.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-all-configs > gpurun_out/r02aj_bench_2gpu.json 2> gpurun_out/r02aj_bench_2gpu.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02aj_bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('workload'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-all-configs --no-eval --config bsds > gpurun_out/r02aj_bench_2gpu_bsds.json 2> gpurun_out/r02aj_bench_2gpu_bsds.err
echo "bench bsds rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02aj_bench_2gpu_bsds.json')); print(d['value'], d['ms_per_step'])"
timeout 200 python bench.py --config hepmass --steps 20 --warmup 5 --no-all-configs --no-cpu-baseline --no-eval > gpurun_out/r02aj_bench_1gpu_hepmass.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02aj_bench_1gpu_hepmass.json')); print('1gpu hepmass', d['value'], d['ms_per_step'])"
