#!/bin/bash
mkdir -p gpurun_out
for v in 0 30 31 32 33; do echo "== variant $v"; RSX_SCATTER_VARIANT=$v SWEEP_FROM=100000 timeout 300 python tools/sweep.py 2>&1 | cut -c1-140 | tail -6; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "histogram_column or trivial or golden" 2>&1 | tail -3
