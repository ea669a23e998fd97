#!/bin/bash
# flakiness check: the GPU suite three times in a row
mkdir -p gpurun_out
for i in 1 2 3; do
  (timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -2) > gpurun_out/r2c62_tests_$i.log; tail -1 gpurun_out/r2c62_tests_$i.log
done
