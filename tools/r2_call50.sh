#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "multi_scale or msg or loss or boundary or sanit" > gpurun_out/r2c50_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c50_tests.log
timeout 200 python tools/msg_bench.py > gpurun_out/r2c50_msg_bench.txt 2>&1; cat gpurun_out/r2c50_msg_bench.txt
