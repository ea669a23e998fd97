#!/bin/bash
# Development aid: per-layer sweep of halo-kernel configurations (RAMNET_HALO_FORCE="PTX,PTY,BN,a_stages,b_stages,pair").
for layer in "gru0 OUT" "gru0 RU" "dec1" "dec2" "enc0" "res"; do
  echo "=== $layer (auto)"; timeout 60 python tools/layer_bench.py --only "$layer" --iters 10 2>&1 | grep -v "^#\|^layer\|^sum"
  for cfg in "1,1,64,2,8,1" "2,1,64,2,8,1" "4,1,64,2,8,1" "2,2,64,2,8,1" "1,1,128,2,6,1" "2,1,128,2,6,1" "4,1,128,1,4,1" "1,1,32,2,8,1" "2,1,32,2,8,1" "4,1,32,2,8,1" "2,2,32,2,8,1" "1,1,256,2,4,1" "2,1,256,1,4,1"; do
    r=$(RAMNET_HALO_FORCE=$cfg timeout 60 python tools/layer_bench.py --only "$layer" --iters 10 2>&1 | grep -v "^#\|^layer\|^sum" | awk '{print $(NF-2), $(NF-1)}')
    echo "  $cfg -> $r"
  done
done
