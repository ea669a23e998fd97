#!/bin/bash
mkdir -p gpurun_out
echo "== static"; timeout 300 python tools/gap_bench.py 2>&1 | tail -14
echo "== dynamic"; RAMNET_FORCE_DYNAMIC=1 timeout 300 python tools/gap_bench.py 2>&1 | tail -14
