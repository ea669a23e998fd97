#!/bin/bash
mkdir -p gpurun_out
for v in 1 0.5 0 ; do
  RAMNET_PLAN_WAVES=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/r2c33_bench_$v.json 2> gpurun_out/r2c33_bench_$v.err; echo "bench PLAN_WAVES=$v rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/r2c33_bench_$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity']['max_rel_err'])"
done
RAMNET_PLAN_WAVES=0 RAMNET_ISSUE_MODEL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/r2c33_bench_x.json 2> gpurun_out/r2c33_bench_x.err; echo "bench PLAN_WAVES=0 ISSUE_MODEL=1 rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2c33_bench_x.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity']['max_rel_err'])"
