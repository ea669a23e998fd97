#!/bin/bash
mkdir -p gpurun_out
for cfg in "up64:RAMNET_UPCONV_MAXC=64" "up128:RAMNET_UPCONV_MAXC=128" "up32:RAMNET_UPCONV_MAXC=32" "off:RAMNET_UPCONV=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  (env $envs RAMNET_DEBUG=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline 2> gpurun_out/r2c7_bench_$name.err) > gpurun_out/r2c7_bench_$name.json
  python -c "
import json
d=json.loads(open('gpurun_out/r2c7_bench_$name.json').read().strip().splitlines()[-1]); r=d['roofline']
print('$name', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],4), 'tf32frac', round(r['frac_of_tf32_pipe'],4), 'conv_ms', round(r['ms_per_step_in_kernel'],2), 'other', {k: round(v,2) for k,v in r['other_kernels_ms_per_step'].items()}, 'parity', d['parity']['max_rel_err'])"
  grep "upconv plan" gpurun_out/r2c7_bench_$name.err | sort -u
done
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/r2c7_tests.log
tail -6 gpurun_out/r2c7_tests.log
