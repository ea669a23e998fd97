#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/r2c3_tests.log
tail -15 gpurun_out/r2c3_tests.log
(timeout 200 python tools/layer_bench.py 2>&1) > gpurun_out/r2c3_layers_base.txt
(RAMNET_HPACK_MAXC=64 timeout 200 python tools/layer_bench.py --only dec 2>&1) > gpurun_out/r2c3_layers_hpack64.txt
for bpc in 25 12; do
 (RAMNET_L2_BPC=$bpc RAMNET_DEBUG=1 timeout 200 python tools/layer_bench.py 2>&1 | grep -v "^\[ramnet\] halo plan.*" ; RAMNET_L2_BPC=$bpc RAMNET_DEBUG=1 timeout 100 python tools/layer_bench.py --iters 1 2>&1 | grep "halo plan" | sort -u) > gpurun_out/r2c3_layers_l2bpc$bpc.txt
done
(RAMNET_PROF=1 RAMNET_DEBUG=1 timeout 200 python tools/layer_bench.py --iters 1 2>&1) > gpurun_out/r2c3_layers_prof.txt
cat gpurun_out/r2c3_layers_base.txt gpurun_out/r2c3_layers_hpack64.txt
