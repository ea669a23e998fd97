#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu -k "gradients_match_reference and fp32 and (bn_train or in_train)" 2>&1 | grep "^E               assert np.float32\|passed\|failed" | cut -c1-120
done
echo "== single stream"
for i in 1 2; do
RAMNET_FRONT_STREAM=0 RAMNET_WGRAD_STREAM=0 timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu -k "gradients_match_reference and fp32 and (bn_train or in_train)" 2>&1 | grep "^E               assert np.float32\|passed\|failed" | cut -c1-120
done
echo "== wgrad stream off only"
RAMNET_WGRAD_STREAM=0 timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu -k "gradients_match_reference and fp32 and (bn_train or in_train)" 2>&1 | grep "^E               assert np.float32\|passed\|failed" | cut -c1-120
echo "== front stream off only"
RAMNET_FRONT_STREAM=0 timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu -k "gradients_match_reference and fp32 and (bn_train or in_train)" 2>&1 | grep "^E               assert np.float32\|passed\|failed" | cut -c1-120
