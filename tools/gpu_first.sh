#!/bin/bash
# First GPU round: parity tests, smoke, bench, yardsticks.  Everything logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 > gpurun_out/cpu.txt; nproc >> gpurun_out/cpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_u32.json 2> gpurun_out/bench_u32.err; tail -c 3000 gpurun_out/bench_u32.json; tail -5 gpurun_out/bench_u32.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload 1B-u64-uniform --no-e2e --no-cpu > gpurun_out/bench_u64.json 2> gpurun_out/bench_u64.err; tail -c 1500 gpurun_out/bench_u64.json; tail -5 gpurun_out/bench_u64.err
timeout 600 ./tools/yardstick 1000000000 > gpurun_out/yardstick.log 2>&1; cat gpurun_out/yardstick.log
