#!/bin/bash
mkdir -p gpurun_out
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_ncu_all_kernels_v3.csv python tools/ncu_all_kernels.py > gpurun_out/r2_ncu_all_kernels_v3.log 2>&1)
python tools/ncu_kernel_table.py gpurun_out/r2_ncu_all_kernels_v3.csv --skip-first-half > gpurun_out/r2_ncu_all_kernels_v3.txt 2>&1
grep -E "relu_bwd|gru_out_bwd|gru_ru_bwd|wgrad|^#" gpurun_out/r2_ncu_all_kernels_v3.txt
(timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2c14_bench.err) > gpurun_out/r2c14_bench.json
python -c "
import json
d=json.loads(open('gpurun_out/r2c14_bench.json').read().strip().splitlines()[-1]); r=d['roofline']; t=d['train']
print('fwd', round(d['value'],1), 'ms', round(d['ms_per_step'],3))
print('train', round(t['maps_per_s'],1), 'ms', round(t['ms_per_step'],2), 'loss', t['loss'], 'launches', t['launches'], t['conv_fwd_dgrad'], t['wgrad']['achieved_tflops'], t['wgrad']['ms_per_step_in_kernel'], t['other_kernels_ms_per_step'])"
