#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c48_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c48_tests.log
timeout 300 python tools/norm_bench.py > gpurun_out/r2c48_norm_bench.txt 2>&1; cat gpurun_out/r2c48_norm_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c48_ncu_norm.csv python tools/norm_bench.py > /dev/null 2>&1; echo "ncu rc=$?"
python tools/ncu_kernel_table.py gpurun_out/r2c48_ncu_norm.csv > gpurun_out/r2c48_ncu_norm.txt 2>&1; head -14 gpurun_out/r2c48_ncu_norm.txt
