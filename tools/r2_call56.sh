#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3) > gpurun_out/r2c56_tests.log; cat gpurun_out/r2c56_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c56_bench.json 2> gpurun_out/r2c56_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2c56_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['parity']['max_rel_err'],d['train']['ms_per_step'],d['gpu_launches'],d['clocks'])"
