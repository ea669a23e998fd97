#!/bin/bash
# live-norm tests again, norm microbench + its ncu launch list, fresh per-layer table, default bench (forward + train leg)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -q -m gpu -s -k "norm or bn_ or in_ or gradients_match_reference" > gpurun_out/r2c42_tests.log 2>&1; echo "tests rc=$?"; grep "vs reference\|passed\|failed" gpurun_out/r2c42_tests.log | tail -8
timeout 300 python tools/norm_bench.py > gpurun_out/r2c42_norm_bench.txt 2>&1; cat gpurun_out/r2c42_norm_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c42_ncu_norm.csv python tools/norm_bench.py > /dev/null 2>&1; echo "ncu rc=$?"
python tools/ncu_kernel_table.py gpurun_out/r2c42_ncu_norm.csv > gpurun_out/r2c42_ncu_norm.txt 2>&1; head -20 gpurun_out/r2c42_ncu_norm.txt
timeout 300 python tools/layer_bench.py > gpurun_out/r2c42_layers.txt 2>&1; grep -v "^\[ramnet" gpurun_out/r2c42_layers.txt | tail -40
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c42_bench.json 2> gpurun_out/r2c42_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c42_bench.err | cut -c1-300
python -c "
import json;d=json.loads(open('gpurun_out/r2c42_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roofline'].get('frac_of_tf32_pipe'),d['parity']); t=d['train']; print({k:t[k] for k in t if k in ('value','ms_per_step','allreduce_ms','loss')})"
