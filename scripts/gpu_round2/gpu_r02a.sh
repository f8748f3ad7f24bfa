#!/bin/bash
# round 2, first GPU call: full GPU test suite (no -x: every failure is wanted), cond-LL parity diagnostics, smoke, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02a_pytest.txt
timeout 900 python scripts/condll_parity.py > gpurun_out/r02a_condll_parity.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -5 gpurun_out/r02a_pytest.txt; tail -3 gpurun_out/r02a_smoke.txt; tail -c 600 gpurun_out/r02a_bench.err; head -c 1500 gpurun_out/r02a_bench.json
