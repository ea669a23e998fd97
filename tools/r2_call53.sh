#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "unet or pred" > gpurun_out/r2c53_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2c53_tests.log
