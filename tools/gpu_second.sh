#!/bin/bash
mkdir -p gpurun_out
./tools/probe_atoms > gpurun_out/probe_atoms.log 2>&1; cat gpurun_out/probe_atoms.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for rm in 0 1; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --rank-mode $rm > gpurun_out/bench_u32_rm$rm.json 2> gpurun_out/bench_u32_rm$rm.err; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/bench_u32_rm$rm.json")); r=j["roofline"]
    print("u32 rank_mode", j["config"]["rank_mode"], "ms", round(j["ms_per_step"],3), "Gkeys/s", round(j["value"],2), "pass ms", round(r["ms_per_launch"],3), "frac", round(r["frac"],3), "hist ms", round(r["histogram_kernel"]["ms"],3))
except Exception as e: print("ERR", e, open("gpurun_out/bench_u32_rm$rm.err").read()[-2000:])
PY
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload 1B-u64-uniform > gpurun_out/bench_u64.json 2> gpurun_out/bench_u64.err; python - <<PY
import json
try:
    j=json.load(open("gpurun_out/bench_u64.json")); r=j["roofline"]
    print("u64 rank_mode", j["config"]["rank_mode"], "ms", round(j["ms_per_step"],3), "Gkeys/s", round(j["value"],2), "pass ms", round(r["ms_per_launch"],3), "frac", round(r["frac"],3), "hist ms", round(r["histogram_kernel"]["ms"],3))
except Exception as e: print("ERR", e, open("gpurun_out/bench_u64.err").read()[-2000:])
PY
# ncu: launch list (shares) and one full capture of the scatter kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload 256M-u32-uniform > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 4 -c 2 -o gpurun_out/prof_scatter_r1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --workload 256M-u32-uniform > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:histogram_kernel -s 1 -c 1 -o gpurun_out/prof_hist_r1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --workload 256M-u32-uniform > gpurun_out/ncu_full_hist.log 2>&1
tail -3 gpurun_out/ncu_full_hist.log
ls -la gpurun_out
