#!/bin/bash
# Round-2 check run (1 GPU): full GPU test-suite, compute-sanitizer on the fused partition/exchange
# kernels, default bench line + reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for tool in memcheck racecheck; do
  echo "== $tool: fused partition + exchange tests"
  timeout 900 compute-sanitizer --tool $tool --target-processes all python -m pytest tests/test_gpu_parity.py -x -q \
     -k "test_scatter_pass_to_destinations or test_key_range_routing" > gpurun_out/sanitize_fused_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_fused_$tool.log | tail -5
done
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 1500 gpurun_out/bench_default.json; echo; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2>&1; tail -c 600 gpurun_out/bench_reference.json; echo
