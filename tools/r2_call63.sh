#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c63_ncu_all.csv python tools/ncu_all_kernels.py > gpurun_out/r2c63_ncu_all.log 2>&1; echo "ncu all rc=$?"; tail -2 gpurun_out/r2c63_ncu_all.log
python tools/ncu_kernel_table.py gpurun_out/r2c63_ncu_all.csv > gpurun_out/r2c63_ncu_all.txt 2>&1; grep -c . gpurun_out/r2c63_ncu_all.txt; grep "norm_\|voxel_grid_fused" gpurun_out/r2c63_ncu_all.txt
