#!/bin/bash
# usage: gpu_r2_sweep.sh "<variants>" "<pytest variants>" "<phase-timing variants>"
mkdir -p gpurun_out
timeout 900 python tools/sweep_variants.py 1000000000 $1 2>&1 | tee gpurun_out/variants.log | cut -c1-220
for v in $2; do
  echo "== pytest under RSX_SCATTER_VARIANT=$v"
  RSX_SCATTER_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "not 1B and not 2_pow" 2>&1 | tail -4
done
for v in $3; do
  for kb in 4 8; do timeout 300 python tools/phase_timing.py 1000000000 $v $kb 2>&1 | tail -14; done
done
