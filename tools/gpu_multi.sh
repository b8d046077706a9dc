#!/bin/bash
# usage: gpu_multi.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
run() { tag=$1; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/bench_multi_${N}_$tag.json 2> gpurun_out/bench_multi_${N}_$tag.err
python - gpurun_out/bench_multi_${N}_$tag.json <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    def line(name, r): print(name, "Gkeys/s", round(r["value"],2), "ms", round(r["ms_per_step"],2), "verified", r["config"]["verified"], "imb", round(r["config"]["imbalance_max_over_mean"],4), {k:(round(v*1e3,2) if isinstance(v,float) else v) for k,v in r["config"]["phase_seconds_rank0"].items()})
    line("u32 1B/GPU", j)
    if "config5_u64" in j: line("u64 2B/GPU", j["config5_u64"])
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-3000:])
PY
}
run fused "$@"
run nccl --no-fused "$@"
