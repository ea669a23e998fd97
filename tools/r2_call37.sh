#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c37_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c37_tests.log
for v in 1 0; do
  RAMNET_FRONT_STREAM=$v timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c37_bench_$v.json 2> gpurun_out/r2c37_bench_$v.err; echo "bench FRONT_STREAM=$v rc=$?"; tail -2 gpurun_out/r2c37_bench_$v.err | cut -c1-300
  python -c "
import json;d=json.loads(open('gpurun_out/r2c37_bench_$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step']); t=d['train']; print({k:t[k] for k in t if k in ('value','ms_per_step','allreduce_ms','loss')})"
done
