#!/bin/bash
# full GPU suite after the live-norm work
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s -k "live_norm" > gpurun_out/r2c41_live_norm.log 2>&1; echo "live norm rc=$?"; grep "max rel err\|passed\|failed" gpurun_out/r2c41_live_norm.log | tail -12
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c41_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2c41_tests.log
