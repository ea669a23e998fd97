#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:halo -s 6 -c 1 -o gpurun_out/r2_ncu_gru0ru -f python tools/layer_bench.py --only "gru0 RU" --iters 4 > gpurun_out/r2c10_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:halo -s 6 -c 1 -o gpurun_out/r2_ncu_res -f python tools/layer_bench.py --only "res 3x3" --iters 4 > gpurun_out/r2c10_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:halo -s 6 -c 1 -o gpurun_out/r2_ncu_dec2up -f python tools/layer_bench.py --only "dec2 upconv 64->32 +pred" --iters 4 > gpurun_out/r2c10_c.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2c10_a.log
