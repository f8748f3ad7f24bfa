#!/bin/bash
# bench.py over every UCI config + the MNIST-config step; one summary line each into gpurun_out/${TAG}_all_configs.txt
set -u
mkdir -p gpurun_out
TAG=${1:-all}
OUT=gpurun_out/${TAG}_all_configs.txt
: > $OUT
for c in gas power hepmass bsds; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_$c.json 2> gpurun_out/${TAG}_$c.err || { echo "$c failed" >> $OUT; continue; }
  python - $c gpurun_out/${TAG}_$c.json >> $OUT <<'PY'
import json, sys
c, d = sys.argv[1], json.load(open(sys.argv[2]))
r, e, b = d["roofline"], d.get("cond_ll_eval") or {}, d.get("reference_batch") or {}
print(f"{c} train {d['value']/1e6:.1f} M/s {d['ms_per_step']:.3f} ms step {r['step_tflops_per_gpu']:.0f} TF ({100*r['step_frac_of_sustained_peak']:.1f}% sust) "
      f"e2e {d['e2e']['value']/1e6:.1f} M/s condLL {e.get('value',0)/1e3:.0f} K/s {e.get('tflops_per_gpu',0):.0f} TF B512 {b.get('ms_per_step',0):.3f} ms")
PY
done
timeout 200 python scripts/mnist_bench.py 2>&1 | tail -1 >> $OUT
cat $OUT
