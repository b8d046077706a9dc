#!/bin/bash
# Round-end evidence run (1 GPU): parity tests, smoke, bench lines, the n-sweep, per-config table,
# ncu launch list + full captures of the dominant kernels.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 400 gpurun_out/bench_default.json; echo
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2>&1; tail -c 300 gpurun_out/bench_reference.json; echo
timeout 600 python bench.py --sweep > gpurun_out/sweep.json 2> gpurun_out/sweep.err; tail -c 300 gpurun_out/sweep.json; echo
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; tail -3 gpurun_out/configs.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra --workload 256M-u32-uniform > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 4 -c 1 -o gpurun_out/prof_scatter_u32 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extra --workload 256M-u32-uniform > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:histogram_kernel -s 1 -c 1 -o gpurun_out/prof_hist_u32 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extra --workload 256M-u32-uniform > gpurun_out/ncu_full_hist.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter2_kernel -s 8 -c 1 -o gpurun_out/prof_scatter_u64 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extra --workload 256M-u64-uniform > gpurun_out/ncu_full64.log 2>&1
ls -la gpurun_out/*.ncu-rep
