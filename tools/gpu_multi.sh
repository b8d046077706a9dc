#!/bin/bash
# usage: gpu_multi.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_multi_$N.json 2> gpurun_out/bench_multi_$N.err
tail -c 3500 gpurun_out/bench_multi_$N.json; tail -5 gpurun_out/bench_multi_$N.err
