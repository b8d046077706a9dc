#!/bin/bash
mkdir -p gpurun_out
echo "== product build"; timeout 900 python tools/sweep_variants.py 1000000000 0 1 2 3 30 2>&1 | tee gpurun_out/variants.log | cut -c1-200
echo "== -DRSX_TICKET_BRANCH build"; RSX_LIB=$PWD/tools/dbg/librsx_branch.so timeout 900 python tools/sweep_variants.py 1000000000 0 2>&1 | cut -c1-200
echo "== product build again"; timeout 900 python tools/sweep_variants.py 1000000000 0 2>&1 | cut -c1-200
