#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2c18_tests.log
tail -15 gpurun_out/r2c18_tests.log
