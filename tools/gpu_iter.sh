#!/bin/bash
# Iteration run: parity tests, bench (u32, u64), optional ncu capture.  Usage: gpu_iter.sh [ncu]
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    j=json.load(open(f)); r=j["roofline"]
    print(j["config"]["workload"], j["config"].get("rank_mode"), "ms", round(j["ms_per_step"],3), "Gkeys/s", round(j["value"],2), "| pass ms", round(r["ms_per_launch"],3), "GB/s", round(r["achieved"]), "frac", round(r["frac"],3), "| hist ms", round(r["histogram_kernel"]["ms"],3), "frac", round(r["histogram_kernel"]["frac"],3), "| sort frac", round(r["whole_sort"]["frac"],3))
except Exception as e: print("ERR", f, e, open(f.replace(".json",".err")).read()[-1500:])
PY
}
for wl in 1B-u32-uniform 1B-u64-uniform 40M-u32-uniform 1B-u32-mask24 500M-f32; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; show gpurun_out/bench_$wl.json
done
if [ "$1" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 4 -c 1 -o gpurun_out/prof_scatter python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --workload 256M-u32-uniform > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
fi
