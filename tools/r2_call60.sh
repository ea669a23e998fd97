#!/bin/bash
mkdir -p gpurun_out
for v in front back 0; do
  RAMNET_DYNAMIC=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline --no-parity > gpurun_out/r2c60_bench_$v.json 2> gpurun_out/r2c60_bench_$v.err; echo "bench DYNAMIC=$v rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/r2c60_bench_$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'])"
done
