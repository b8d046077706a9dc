#!/bin/bash
# Round-end evidence run (1 GPU): parity tests, smoke, bench lines, ncu launch list + full captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json; echo
timeout 900 python bench.py --workload 1B-u64-uniform --steps 5 > gpurun_out/bench_u64.json 2> gpurun_out/bench_u64.err; tail -c 300 gpurun_out/bench_u64.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>&1; tail -c 400 gpurun_out/bench_reference.json; echo
timeout 600 python tools/sweep.py > gpurun_out/sweep.log 2>&1; tail -12 gpurun_out/sweep.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload 256M-u32-uniform > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 4 -c 1 -o gpurun_out/prof_scatter_final python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --workload 256M-u32-uniform > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:histogram_kernel -s 1 -c 1 -o gpurun_out/prof_hist_final python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --workload 256M-u32-uniform > gpurun_out/ncu_full_hist.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 8 -c 1 -o gpurun_out/prof_scatter_final_u64 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --workload 256M-u64-uniform > gpurun_out/ncu_full64.log 2>&1
ls gpurun_out/*.ncu-rep
