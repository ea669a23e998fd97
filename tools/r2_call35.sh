#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2c35_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c35_tests.log
for v in 1 0; do
  RAMNET_WGRAD_STREAM=$v timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c35_bench_$v.json 2> gpurun_out/r2c35_bench_$v.err; echo "bench WGRAD_STREAM=$v rc=$?"; tail -2 gpurun_out/r2c35_bench_$v.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2c35_bench_$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step']); t=d['train']; print({k:t[k] for k in t if k in ('value','ms_per_step','allreduce_ms','loss')})"
done
