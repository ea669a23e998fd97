#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_experimental.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r2c25_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c25_tests.log
RAMNET_DEBUG=1 timeout 300 python tools/layer_bench.py --only enc 2>&1 | grep -E "enc|s2seg plan" | sort | uniq | head -20
for v in "" "RAMNET_ISSUE_MODEL=1"; do
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/r2c25_bench.json 2> gpurun_out/r2c25_bench.err; echo "bench [$v] rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/r2c25_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['parity']['max_rel_err'])"
done
env RAMNET_ISSUE_MODEL=1 timeout 300 python tools/layer_bench.py 2>&1 | tail -20
