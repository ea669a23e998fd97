#!/bin/bash
mkdir -p gpurun_out
for v in 0 0.5; do
  RAMNET_PLAN_WAVES=$v timeout 900 python bench.py --steps 5 --warmup 3 --no-parity > gpurun_out/r2c36_bench_$v.json 2> gpurun_out/r2c36_bench_$v.err; echo "bench PLAN_WAVES=$v rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/r2c36_bench_$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step']); t=d['train']; print({k:t[k] for k in t if k in ('value','ms_per_step','allreduce_ms','loss')})"
done
