#!/bin/bash
# A/B runs of scatter tuning variants (rsx_scatter.cuh).  Usage: gpurun -- bash tools/gpu_ab.sh "<variants>"
mkdir -p gpurun_out
VARS=${1:-"0 1 2"}
for wl in 1B-u32-uniform 1B-u64-uniform; do
  for v in $VARS; do
    timeout 120 python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu --variant $v > gpurun_out/ab_${wl}_${v}.json 2> gpurun_out/ab_${wl}_${v}.err || tail -5 gpurun_out/ab_${wl}_${v}.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${wl}_${v}.json").read().strip().splitlines()[-1])
    print("$wl variant=$v", round(d["ms_per_step"],3), "ms", round(d["value"],2), "Gk/s pass", round(d["roofline"]["ms_per_launch"],3), "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("$wl variant=$v FAILED", e)
PY
  done
done
