#!/bin/bash
# Development aid: per-layer counters of the tap-packed wgrad kernel for a few planner settings (RAMNET_WGP="TR,RG,CTAs/SM,smem KB").
for cfg in ${WGP_CFGS:-"8,3,2,200" "8,2,4,100" "4,3,2,200"}; do
  echo "=== RAMNET_WGP=$cfg"
  RAMNET_WGP=$cfg RAMNET_PROF=1 timeout 120 python tools/wgrad_bench.py --iters 1 2>&1 | grep "wgrad packed" | awk 'NR%3==0' | sed 's/\[ramnet-prof\] wgrad packed //'
done
