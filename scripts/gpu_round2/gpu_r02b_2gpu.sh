#!/bin/bash
# round 2, 2-GPU call: the 2-rank Trainer test (overlapped bucketed all-reduce inside one CUDA graph vs serial vs 1 rank)
# and the bench at N = 2 with and without the overlap
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r02b_pytest_dist.txt
for mode in 1 0; do
  PMVAE_DP_OVERLAP=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-all-configs --no-eval > gpurun_out/r02b_bench_2gpu_overlap$mode.json 2> gpurun_out/r02b_bench_2gpu_overlap$mode.err
done
timeout 600 python bench.py --config hepmass --steps 20 --warmup 5 --no-all-configs --no-eval --no-cpu-baseline > gpurun_out/r02b_bench_1gpu_hepmass.json 2> gpurun_out/r02b_bench_1gpu_hepmass.err
tail -8 gpurun_out/r02b_pytest_dist.txt
for f in gpurun_out/r02b_bench_*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; done
tail -c 1500 gpurun_out/r02b_bench_2gpu_overlap1.err
