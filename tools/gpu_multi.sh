#!/bin/bash
# multi-GPU check: C ABI tests, the CLI, torchrun bench at N = number of visible GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 300 ./tools/radix_multi_b200 1000000000 $N uint32_t 2>&1 | tail -2
timeout 300 ./tools/radix_multi_b200 1000000000 $N uint64_t zipf 2>&1 | tail -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_multi_$N.json 2> gpurun_out/bench_multi_$N.err
tail -c 3000 gpurun_out/bench_multi_$N.json; echo; grep -v "^W\|^\[W\|warn" gpurun_out/bench_multi_$N.err | tail -15
