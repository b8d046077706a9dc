#!/bin/bash
# compute-sanitizer on small runs of the real entry points (C++ drop-in programs)
mkdir -p gpurun_out; cd tools
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool: radix_tests_b200"; timeout 600 compute-sanitizer --tool $tool ./radix_tests_b200 2>&1 | tail -4
  echo "== $tool: radix_b200 300000 u32"; timeout 600 compute-sanitizer --tool $tool ./radix_b200 300000 0 0 uint32_t 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Sorted|hazard|error" | head -6
  echo "== $tool: radix_b200 200000 u64"; timeout 600 compute-sanitizer --tool $tool ./radix_b200 200000 0 0 uint64_t 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Sorted|hazard|error" | head -6
done
