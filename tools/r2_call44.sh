#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/norm_bench.py > gpurun_out/r2c44_norm_bench.txt 2>&1; cat gpurun_out/r2c44_norm_bench.txt
timeout 300 python tools/gap_bench.py > gpurun_out/r2c44_gap.txt 2>&1; cat gpurun_out/r2c44_gap.txt
RAMNET_PDL=1 timeout 300 python tools/gap_bench.py > gpurun_out/r2c44_gap_pdl.txt 2>&1; echo "== PDL"; cat gpurun_out/r2c44_gap_pdl.txt
RAMNET_PROF=1 RAMNET_DEBUG=1 timeout 300 python tools/layer_bench.py --iters 2 > gpurun_out/r2c44_layers_prof.txt 2>&1; grep -v "^\[ramnet\] halo plan" gpurun_out/r2c44_layers_prof.txt | awk '!seen[$0]++' | head -80
