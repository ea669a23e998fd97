#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_experimental.py -x -q -m gpu -k s2seg > gpurun_out/r2c23_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c23_tests.log
for L in enc0 enc1 enc2; do
for f in 1,1,64,0 1,1,64,1 2,1,64,0 2,1,64,1 2,2,64,1 4,1,64,1 1,1,128,0 1,1,128,1 2,1,128,1 2,2,128,1 2,1,128,0 1,1,256,1 2,1,256,1; do
  echo -n "$L force=$f: "
  RAMNET_S2_FORCE=$f timeout 100 python tools/layer_bench.py --only $L 2>&1 | grep -E "^$L" | awk '{print $(NF-2), $(NF-1)}'
done; done
