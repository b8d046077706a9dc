#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --no-e2e --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
for k,r in (("u32",d),("u64",d["u64_1B"])):
    print(k, "ms", round(r["ms_per_step"],3), "Gk/s", round(r["value"],2), "pass ms", round(r["roofline"]["ms_per_launch"],3), "frac", round(r["roofline"]["frac"],3), "hist ms", round(r["roofline"]["histogram_kernel"]["ms"],3), "launches", r["gpu_launches"])
PY
tail -3 gpurun_out/bench_quick.err
timeout 300 python tools/sweep.py 2>&1 | tail -12
