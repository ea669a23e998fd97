#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_train.py -m gpu -q -k "inference_output or transposed" 2>&1 | grep -v "^$" | tail -60) > gpurun_out/r2c19_tests.log
grep -E "^E |passed|failed" gpurun_out/r2c19_tests.log | cut -c1-300 | head -30
