#!/bin/bash
for L in "gru0 RU" "gru0 OUT" "gru1 RU"; do
for v in "RAMNET_AUX_TMA=0" "RAMNET_AUX_TMA=1" "RAMNET_AUX_TMA=1 RAMNET_AUX_SLOTS=2"; do
  echo "== $L $v"
  env $v RAMNET_PROF=1 RAMNET_DEBUG=1 timeout 100 python tools/layer_bench.py --only "$L" --iters 1 2>&1 | grep -E "halo plan|ramnet-prof" | tail -2 | cut -c1-260
done; done
