#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3) > gpurun_out/r2c52_tests.log; cat gpurun_out/r2c52_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c52_smoke.txt 2>&1; tail -1 gpurun_out/r2c52_smoke.txt
timeout 900 python bench.py > gpurun_out/r2c52_bench.json 2> gpurun_out/r2c52_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2c52_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roofline'].get('frac_of_tf32_pipe'),d['parity']['max_rel_err'],d['cpu_baseline']['value'],d['roofline']['traffic']); t=d['train']; print({k:t[k] for k in t if k in ('maps_per_s','ms_per_step','loss')})"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c52_ncu_all.csv python tools/ncu_all_kernels.py > gpurun_out/r2c52_ncu_all.log 2>&1; echo "ncu all rc=$?"
python tools/ncu_kernel_table.py gpurun_out/r2c52_ncu_all.csv --skip-first-half > gpurun_out/r2c52_ncu_all.txt 2>&1; grep -c . gpurun_out/r2c52_ncu_all.txt; grep "msg_\|voxel" gpurun_out/r2c52_ncu_all.txt
