#!/bin/bash
mkdir -p gpurun_out
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_ncu_all_kernels_v2.csv python tools/ncu_all_kernels.py > gpurun_out/r2_ncu_all_kernels_v2.log 2>&1)
tail -2 gpurun_out/r2_ncu_all_kernels_v2.log
python tools/ncu_kernel_table.py gpurun_out/r2_ncu_all_kernels_v2.csv --skip-first-half > gpurun_out/r2_ncu_all_kernels_v2.txt 2>&1
head -45 gpurun_out/r2_ncu_all_kernels_v2.txt
