#!/bin/bash
# Round 2, GPU call 2: full GPU test suite, default bench (parity gate + train leg), sanitizers, every-kernel ncu list.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2c2_tests.log
tail -5 gpurun_out/r2c2_tests.log
(timeout 600 python bench.py 2> gpurun_out/r2c2_bench.err) > gpurun_out/r2c2_bench.json
echo "bench rc=$?"; cut -c1-400 gpurun_out/r2c2_bench.json; tail -5 gpurun_out/r2c2_bench.err
bash tools/r2_sanitize.sh
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_ncu_all_kernels.csv python tools/ncu_all_kernels.py > gpurun_out/r2_ncu_all_kernels.log 2>&1)
tail -2 gpurun_out/r2_ncu_all_kernels.log
python tools/ncu_kernel_table.py gpurun_out/r2_ncu_all_kernels.csv --skip-first-half > gpurun_out/r2_ncu_all_kernels.txt 2>&1
head -40 gpurun_out/r2_ncu_all_kernels.txt
