#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/layer_bench.py 2>&1) > gpurun_out/r2c11_layers.txt
(timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline 2> gpurun_out/r2c11_bench.err) > gpurun_out/r2c11_bench.json
python -c "
import json
d=json.loads(open('gpurun_out/r2c11_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('bench', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],4), 'tf32', round(r['frac_of_tf32_pipe'],4), 'conv_ms', round(r['ms_per_step_in_kernel'],2), r['other_kernels_ms_per_step'], 'parity', d['parity']['max_rel_err'])"
(timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_kernels.py tests/test_gpu_upconv.py -m gpu -q 2>&1 | tail -5) > gpurun_out/r2c11_tests.log
cat gpurun_out/r2c11_layers.txt; tail -3 gpurun_out/r2c11_tests.log
