#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r2c34_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c34_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-train > gpurun_out/r2c34_bench.json 2> gpurun_out/r2c34_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2c34_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['parity']['max_rel_err'], d['roofline']['frac'])"
