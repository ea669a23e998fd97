#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_upconv.py -m gpu -q -x 2>&1 | tail -40) > gpurun_out/r2c6_upconv_tests.log
tail -25 gpurun_out/r2c6_upconv_tests.log
