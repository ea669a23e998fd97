#!/bin/bash
mkdir -p gpurun_out
for cfg in "up64:RAMNET_UPCONV_MAXC=64" "up32:RAMNET_UPCONV_MAXC=32"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  (env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline 2> gpurun_out/r2c9_bench_$name.err) > gpurun_out/r2c9_bench_$name.json
  python -c "
import json
d=json.loads(open('gpurun_out/r2c9_bench_$name.json').read().strip().splitlines()[-1]); r=d['roofline']
print('$name', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(r['frac'],4), 'step frac', round(r['frac_of_step'],4), 'parity', d['parity']['max_rel_err'])"
done
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2c9_tests.log
tail -4 gpurun_out/r2c9_tests.log
