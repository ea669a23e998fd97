#!/bin/bash
for L in "gru0 RU" "res" "dec0 5x5" "gru2 OUT"; do
  echo "== $L"
  RAMNET_PROF=1 timeout 100 python tools/layer_bench.py --only "$L" --iters 1 2>&1 | grep -E "ramnet-prof" | tail -2 | cut -c1-300
  timeout 100 python tools/layer_bench.py --only "$L" 2>&1 | grep -E "^(gru|res|dec)"
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
