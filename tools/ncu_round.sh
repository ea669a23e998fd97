#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the conv / wgrad kernels (run under gpurun).
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches_v7.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_v7.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:halo -s 60 -c 20 -o gpurun_out/halo_pass_v7 -f \
    python bench.py --no-graphs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_halo_v7.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:wgrad_packed_kernel -s 20 -c 10 -o gpurun_out/wgrad_packed_v3 -f \
    python bench.py --mode train --no-graphs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_wgrad_v3.log 2>&1
ls -la gpurun_out/*.ncu-rep
