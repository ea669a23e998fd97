#!/bin/bash
for cfg in "enc0 1,1,64,1" "enc0 2,1,64,1" "enc0 2,2,64,1" "enc1 2,1,128,1" "enc1 1,1,128,1" "enc2 1,1,128,1" "enc2 2,1,64,1"; do
  set -- $cfg
  echo "== $1 force=$2"
  RAMNET_PROF=1 RAMNET_DEBUG=1 RAMNET_S2_FORCE=$2 timeout 100 python tools/layer_bench.py --only $1 --iters 1 2>&1 | grep -E "s2seg plan|ramnet-prof" | tail -2
done
