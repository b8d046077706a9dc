#!/bin/bash
# A/B runs of scatter variants.  Usage: gpurun -- bash tools/gpu_ab.sh "<variants>" "<bulk settings>"
mkdir -p gpurun_out
VARS=${1:-"0 1 2"}
BULKS=${2:-"0"}
for wl in 1B-u32-uniform 1B-u64-uniform; do
  for v in $VARS; do
  for b in $BULKS; do
    timeout 120 python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu --bulk-store $b --variant $v > gpurun_out/ab_${wl}_${v}_$b.json 2> gpurun_out/ab_${wl}_${v}_$b.err || tail -5 gpurun_out/ab_${wl}_${v}_$b.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${wl}_${v}_$b.json").read().strip().splitlines()[-1])
    print("$wl variant=$v bulk=$b", round(d["ms_per_step"],3), "ms", round(d["value"],2), "Gk/s pass", round(d["roofline"]["ms_per_launch"],3), "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("$wl variant=$v bulk=$b FAILED", e)
PY
  done
  done
done
