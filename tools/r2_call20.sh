#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_boundary.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -8) > gpurun_out/r2c20_tests.log
tail -4 gpurun_out/r2c20_tests.log
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_ncu_all_kernels_v4.csv python tools/ncu_all_kernels.py > gpurun_out/r2_ncu_all_kernels_v4.log 2>&1)
python tools/ncu_kernel_table.py gpurun_out/r2_ncu_all_kernels_v4.csv --skip-first-half > gpurun_out/r2_ncu_all_kernels_v4.txt 2>&1
grep -E "relu_bwd|gru_out_bwd|gru_ru_bwd|^#" gpurun_out/r2_ncu_all_kernels_v4.txt
(timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2c20_bench.err) > gpurun_out/r2c20_bench.json
python -c "
import json
d=json.loads(open('gpurun_out/r2c20_bench.json').read().strip().splitlines()[-1]); r=d['roofline']; t=d['train']
print('fwd', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],4))
print('train', round(t['maps_per_s'],1), 'ms', round(t['ms_per_step'],2), 'loss', t['loss'], 'launches', t['launches'], t['conv_fwd_dgrad'], t['wgrad']['achieved_tflops'], t['wgrad']['ms_per_step_in_kernel'])"
