#!/bin/bash
# usage: gpu_multi.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
run() { tag=$1; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/bench_multi_${N}_$tag.json 2> gpurun_out/bench_multi_${N}_$tag.err
python tools/show_multi.py gpurun_out/bench_multi_${N}_$tag.json
}
run fused "$@"
run nccl --no-fused "$@"
