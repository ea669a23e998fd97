#!/bin/bash
mkdir -p gpurun_out
(RAMNET_DEBUG=1 timeout 200 python tools/layer_bench.py 2>&1 | sort -u -k1,1 -k2 | grep -v "^$") > gpurun_out/r2c5_layers_new.txt
(timeout 200 python tools/layer_bench.py 2>&1) > gpurun_out/r2c5_layers_new_clean.txt
(RAMNET_EPI_MODEL=0 timeout 200 python tools/layer_bench.py 2>&1) > gpurun_out/r2c5_layers_oldmodel.txt
(RAMNET_HPACK_MAXC=64 timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline 2>/dev/null) > gpurun_out/r2c5_bench_hp64.json
(RAMNET_HPACK_MAXC=64 RAMNET_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline 2>/dev/null) > gpurun_out/r2c5_bench_hp64_pdl.json
(timeout 300 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline 2>/dev/null) > gpurun_out/r2c5_bench.json
(timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -8) > gpurun_out/r2c5_tests.log
cat gpurun_out/r2c5_layers_new_clean.txt; paste <(cut -c1-60 gpurun_out/r2c5_layers_oldmodel.txt) | tail -15
for f in r2c5_bench r2c5_bench_hp64 r2c5_bench_hp64_pdl; do python -c "
import json,sys
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'parity', d['parity']['max_rel_err'])"; done
tail -4 gpurun_out/r2c5_tests.log
