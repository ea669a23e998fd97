#!/bin/bash
mkdir -p gpurun_out
T='timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu -k "gradients_match_reference and fp32 and (bn_train or in_train)"'
echo "== float64 per element"; eval $T 2>&1 | grep "^E               assert np.float32\|passed\|failed" | cut -c1-120
timeout 300 python tools/norm_bench.py 2>&1 | head -6
for v in 1 2; do
  RAMNET_NVCC_EXTRA="-DRAMNET_NORM_F32_RUNS=$v" python -m rpg_ramnet_b200.build --force > /dev/null 2>&1
  echo "== F32_RUNS=$v"; eval $T 2>&1 | grep "^E               assert np.float32\|passed\|failed" | cut -c1-120
done
