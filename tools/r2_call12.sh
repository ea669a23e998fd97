#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_boundary.py tests/test_gpu_fullsize.py tests/test_gpu_experimental.py -m gpu -q 2>&1 | tail -40) > gpurun_out/r2c12_tests.log
tail -12 gpurun_out/r2c12_tests.log
(timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2c12_bench.err) > gpurun_out/r2c12_bench.json
python -c "
import json
d=json.loads(open('gpurun_out/r2c12_bench.json').read().strip().splitlines()[-1]); r=d['roofline']; t=d['train']
print('fwd', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],4), 'tf32', round(r['frac_of_tf32_pipe'],4), 'conv_ms', round(r['ms_per_step_in_kernel'],2), r['other_kernels_ms_per_step'])
print('train', round(t['maps_per_s'],1), 'ms', round(t['ms_per_step'],2), 'loss', t['loss'], 'launches', t['launches'], t['conv_fwd_dgrad'], t['wgrad']['achieved_tflops'], t['wgrad']['ms_per_step_in_kernel'], t['other_kernels_ms_per_step'])"
tail -3 gpurun_out/r2c12_bench.err
