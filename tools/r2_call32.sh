#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_model.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2c32_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c32_tests.log
for v in 1 0; do
  RAMNET_PASS_OVERLAP=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/r2c32_bench_$v.json 2> gpurun_out/r2c32_bench_$v.err; echo "bench OVERLAP=$v rc=$?"; tail -3 gpurun_out/r2c32_bench_$v.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2c32_bench_$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity']['max_rel_err'])"
done
