#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
RAMNET_AUX_TMA=$v timeout 600 ncu --set full --import-source on --clock-control none -k regex:halo_kernel -s 3 -c 1 -f -o gpurun_out/r2c27_gru0ru_aux$v python tools/layer_bench.py --only "gru0 RU" --iters 2 > gpurun_out/r2c27_ncu_$v.log 2>&1; echo "ncu rc=$?"
done
ls -la gpurun_out/*.ncu-rep
