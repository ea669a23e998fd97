#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_experimental.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r2c26_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c26_tests.log
for v in 1 0; do
  echo "== RAMNET_AUX_TMA=$v"
  RAMNET_AUX_TMA=$v RAMNET_DEBUG=1 timeout 300 python tools/layer_bench.py --only "gru" 2>&1 | grep -E "^gru|halo plan" | sort | uniq | head -20
  RAMNET_AUX_TMA=$v timeout 300 python tools/layer_bench.py --only "res" 2>&1 | grep -E "^res"
done
for v in 1 0; do
  RAMNET_AUX_TMA=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/r2c26_bench_$v.json 2> gpurun_out/r2c26_bench_$v.err; echo "bench AUX_TMA=$v rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/r2c26_bench_$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['parity']['max_rel_err'])"
done
