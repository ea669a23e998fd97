#!/bin/bash
for kb in 40 48; do
echo "== RAMNET_BSTAGE_KB=$kb"
for L in "gru1 RU" "gru2 RU" "gru2 OUT" "res"; do
RAMNET_BSTAGE_KB=$kb timeout 300 python tools/plan_sweep.py "$L" 2>&1 | grep -v "^sum"
done; done
