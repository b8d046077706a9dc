#!/bin/bash
# multi-GPU check: C ABI tests, the CLI, torchrun bench at N = number of visible GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "multi or selftest or scatter_pass_to or key_range" 2>&1 | tail -5
timeout 300 ./tools/radix_multi_b200 1000000000 $N uint32_t 2>&1 | tail -1
timeout 300 ./tools/radix_multi_b200 1000000000 $N uint64_t 2>&1 | tail -1
timeout 300 ./tools/radix_multi_b200 1000000000 $N uint64_t zipf 2>&1 | tail -1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_multi_$N.json 2> gpurun_out/bench_multi_$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_multi_{N}.json").read().strip().splitlines()[-1])
    print("N", N, "Gk/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), d["config"]["phase_seconds_rank0"], "verified", d["config"]["verified"], "selftest", d["selftest"]["passed"] if d.get("selftest") else None)
    for k in ("config5_u64", "config5_u64_zipf"):
        if k in d: print(" ", k, round(d[k]["value"], 1), "Gk/s", round(d[k]["ms_per_step"], 2), "ms", d[k]["config"]["phase_seconds_rank0"], d[k]["config"]["verified"])
    print("  e2e", d["e2e"]["value"] if d.get("e2e") else None)
except Exception as e:
    print("ERR", e); print(open(f"gpurun_out/bench_multi_{N}.err").read()[-2500:])
PY
