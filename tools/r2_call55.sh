#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "graph or unsupported" > gpurun_out/r2c55_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2c55_tests.log
