#!/bin/bash
# tcgen05 bring-up: probe each kernel kind in its own process (a trap kills the context),
# then the full GPU suite and a short bench.
set -u
mkdir -p gpurun_out
TAG=${1:-tc}
timeout 300 python scripts/tc_probe.py nt > gpurun_out/${TAG}_probe_nt.log 2>&1; echo "probe nt exit $?"
tail -25 gpurun_out/${TAG}_probe_nt.log
timeout 300 python scripts/tc_probe.py tn > gpurun_out/${TAG}_probe_tn.log 2>&1; echo "probe tn exit $?"
tail -25 gpurun_out/${TAG}_probe_tn.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
tail -c 2500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
