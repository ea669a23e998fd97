#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k voxel 2>&1 | tail -8) > gpurun_out/r2c16_voxel_tests.log
tail -5 gpurun_out/r2c16_voxel_tests.log
(timeout 300 python tools/voxel_bench.py --json gpurun_out/r2_voxel_bench.json 2>&1) > gpurun_out/r2c16_voxel_bench.txt
(RAMNET_VOXEL_PACKED_MIN=999999999999 timeout 300 python tools/voxel_bench.py 2>&1) > gpurun_out/r2c16_voxel_bench_direct.txt
cat gpurun_out/r2c16_voxel_bench.txt; echo "== direct kernel only"; cat gpurun_out/r2c16_voxel_bench_direct.txt
