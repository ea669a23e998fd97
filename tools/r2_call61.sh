#!/bin/bash
mkdir -p gpurun_out
for v in front back none; do
  RAMNET_STREAM_PRIORITY=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r2c61_bench_$v.json 2> gpurun_out/r2c61_bench_$v.err; echo "bench PRIORITY=$v rc=$?"; tail -1 gpurun_out/r2c61_bench_$v.err | cut -c1-200
  python -c "
import json;d=json.loads(open('gpurun_out/r2c61_bench_$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity']['max_rel_err'],d['clocks'])"
done
