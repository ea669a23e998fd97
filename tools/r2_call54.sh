#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/wgrad_bench.py > gpurun_out/r2c54_wgrad.txt 2>&1; grep -v "^\[ramnet" gpurun_out/r2c54_wgrad.txt | tail -30
timeout 300 python tools/train_profile.py > gpurun_out/r2c54_train_profile.txt 2>&1; tail -45 gpurun_out/r2c54_train_profile.txt
