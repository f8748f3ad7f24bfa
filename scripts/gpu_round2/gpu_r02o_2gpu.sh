#!/bin/bash
# 2 GPUs: the rank-exit fix (no teardown hang), the 2-rank Trainer test, N = 2 bench line (short timeouts everywhere)
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/r02o_pytest_dist.txt
tail -6 gpurun_out/r02o_pytest_dist.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02o_bench_2gpu.json 2> gpurun_out/r02o_bench_2gpu.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02o_bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k: v['train']['value'] for k, v in d['all_configs'].items()})"
tail -c 600 gpurun_out/r02o_bench_2gpu.err
