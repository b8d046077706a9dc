#!/bin/bash
# usage: gpu_multi_final.sh N  -- default bench (fused), NCCL baseline, zipf config-5 variant
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_multi_${N}_fused.json 2> gpurun_out/bench_multi_${N}_fused.err
python tools/show_multi.py gpurun_out/bench_multi_${N}_fused.json | cut -c1-500
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/bench_multi_${N}_fused.json").read().splitlines() if l.startswith("{")][-1]); print("e2e", j["e2e"])
except Exception as e: print(e)
PY
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-fused --no-e2e > gpurun_out/bench_multi_${N}_nccl.json 2> gpurun_out/bench_multi_${N}_nccl.err
python tools/show_multi.py gpurun_out/bench_multi_${N}_nccl.json | cut -c1-500
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --workload 2B-u64-zipf > gpurun_out/bench_multi_${N}_u64zipf.json 2> gpurun_out/bench_multi_${N}_u64zipf.err
python tools/show_multi.py gpurun_out/bench_multi_${N}_u64zipf.json | cut -c1-500
