#!/bin/bash
mkdir -p gpurun_out
(timeout 200 python tools/layer_bench.py --only upconv 2>&1) > gpurun_out/r2c8_up_layers.txt
(RAMNET_PROF=1 RAMNET_DEBUG=1 timeout 200 python tools/layer_bench.py --only upconv --iters 1 2>&1) > gpurun_out/r2c8_up_layers_prof.txt
for f in "2,1,128,1" "1,1,128,1" "2,2,128,1" "4,1,128,1" "2,1,128,0" "1,1,128,0"; do echo "== dec2 FORCE $f"; RAMNET_UP_FORCE=$f RAMNET_DEBUG=1 timeout 100 python tools/layer_bench.py --only "dec2 upconv" 2>&1 | grep -E "upconv" | sort -u; done > gpurun_out/r2c8_up_force.txt 2>&1
for f in "1,1,256,1" "2,1,128,1" "1,1,128,1" "2,1,256,0" "1,1,256,0"; do echo "== dec1 FORCE $f"; RAMNET_UP_FORCE=$f RAMNET_DEBUG=1 timeout 100 python tools/layer_bench.py --only "dec1 upconv" 2>&1 | grep -E "upconv" | sort -u; done >> gpurun_out/r2c8_up_force.txt 2>&1
cat gpurun_out/r2c8_up_layers.txt; grep -E "prof\]|plan|upconv" gpurun_out/r2c8_up_layers_prof.txt | cut -c1-330 | awk '!seen[$0]++' | head -30; cat gpurun_out/r2c8_up_force.txt | cut -c1-200
