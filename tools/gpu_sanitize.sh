#!/bin/bash
# compute-sanitizer on small runs of the real entry points: the C++ drop-in programs and the
# fused partition + exchange / key-range kernels (through their pytest cases).
mkdir -p gpurun_out; 
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool: radix_tests_b200"; (cd tools && timeout 600 compute-sanitizer --tool $tool ./radix_tests_b200 2>&1 | tail -4)
  echo "== $tool: radix_b200 300000 u32"; (cd tools && timeout 600 compute-sanitizer --tool $tool ./radix_b200 300000 0 0 uint32_t 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Sorted|hazard|error" | head -6)
  echo "== $tool: radix_b200 200000 u64"; (cd tools && timeout 600 compute-sanitizer --tool $tool ./radix_b200 200000 0 0 uint64_t 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Sorted|hazard|error" | head -6)
done
for tool in memcheck racecheck; do
  echo "== $tool: fused partition + exchange, key-range routing, key compaction, any-size records (pytest)"
  timeout 1200 compute-sanitizer --tool $tool --target-processes all python -m pytest tests/test_gpu_parity.py -x -q \
     -k "test_scatter_pass_to_destinations or test_key_range_routing or (test_key_compaction and u32) or (any_size and 1000)" > gpurun_out/sanitize_pytest_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_pytest_$tool.log | tail -5
done
