#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "norm or voxel or bn_ or in_" > gpurun_out/r2c45_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c45_tests.log
timeout 300 python tools/norm_bench.py > gpurun_out/r2c45_norm_bench.txt 2>&1; cat gpurun_out/r2c45_norm_bench.txt
timeout 600 python tools/voxel_bench.py --json gpurun_out/r2c45_voxel.json > gpurun_out/r2c45_voxel.txt 2>&1; cat gpurun_out/r2c45_voxel.txt
echo "== unfused"; RAMNET_VOXEL_FUSED=0 timeout 600 python tools/voxel_bench.py > gpurun_out/r2c45_voxel_unfused.txt 2>&1; head -8 gpurun_out/r2c45_voxel_unfused.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c45_ncu_norm.csv python tools/norm_bench.py > /dev/null 2>&1; echo "ncu rc=$?"
python tools/ncu_kernel_table.py gpurun_out/r2c45_ncu_norm.csv > gpurun_out/r2c45_ncu_norm.txt 2>&1; head -16 gpurun_out/r2c45_ncu_norm.txt
