#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "dynamic_work" > gpurun_out/r2c58_dyn_tests.log 2>&1; echo "dyn kernel tests rc=$?"; tail -12 gpurun_out/r2c58_dyn_tests.log
timeout 400 python -m pytest tests/test_gpu_model.py tests/test_gpu_boundary.py -x -q -m gpu -k "graph" > gpurun_out/r2c58_graph_tests.log 2>&1; echo "graph tests rc=$?"; tail -6 gpurun_out/r2c58_graph_tests.log
for v in 0 1 0 1; do
  RAMNET_DYNAMIC=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r2c58_bench_$v.json 2> gpurun_out/r2c58_bench_$v.err; echo "bench DYNAMIC=$v rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/r2c58_bench_$v.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity']['max_rel_err'],d['clocks'])"
done
