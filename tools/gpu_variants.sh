#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    j=json.load(open(f)); r=j["roofline"]
    print(j["config"]["workload"], "v", j["config"].get("variant"), "ms", round(j["ms_per_step"],3), "Gkeys/s", round(j["value"],2), "| pass ms", round(r["ms_per_launch"],3), "GB/s", round(r["achieved"]), "frac", round(r["frac"],3), "| hist ms", round(r["histogram_kernel"]["ms"],3))
except Exception as e: print("ERR", f, e, open(f.replace(".json",".err")).read()[-800:])
PY
}
for wl in $WORKLOADS; do
for v in $VARIANTS; do
timeout 300 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --workload $wl --variant $v > gpurun_out/bench_${wl}_v$v.json 2> gpurun_out/bench_${wl}_v$v.err; show gpurun_out/bench_${wl}_v$v.json
done; done
