#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_train.py tests/test_gpu_kernels.py tests/test_gpu_boundary.py -m gpu -q -s 2>&1 | grep -v "^tests/.*PASSED" | tail -300) > gpurun_out/r2c4_tests.log
grep -E "passed|failed|tf32 |Error|assert" gpurun_out/r2c4_tests.log | head -40
for f in r2c3_layers_l2bpc25 r2c3_layers_l2bpc12; do :; done
