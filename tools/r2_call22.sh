#!/bin/bash
# s2seg: correctness, per-layer A/B, bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_experimental.py -x -q -m gpu -k s2seg > gpurun_out/r2c22_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c22_tests.log
for v in 1 0; do
  echo "== RAMNET_S2SEG=$v"
  RAMNET_S2SEG=$v RAMNET_DEBUG=1 timeout 300 python tools/layer_bench.py --only enc 2>&1 | grep -E "enc|s2seg plan|halo plan s2" | sort | uniq | head -20
done
for v in 1 0; do
  RAMNET_S2SEG=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/r2c22_bench_$v.json 2> gpurun_out/r2c22_bench_$v.err; echo "bench S2SEG=$v rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/r2c22_bench_$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d.get('parity'))"
done
