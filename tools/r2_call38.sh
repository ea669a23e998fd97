#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-parity > gpurun_out/r2c38_bench.json 2> gpurun_out/r2c38_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c38_bench.err | cut -c1-300
python -c "
import json;d=json.loads(open('gpurun_out/r2c38_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step']); t=d['train']; print({k:t[k] for k in t if k in ('value','ms_per_step','allreduce_ms','loss')})"
