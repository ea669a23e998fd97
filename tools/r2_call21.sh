#!/bin/bash
mkdir -p gpurun_out
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_ncu_all_kernels_v5.csv python tools/ncu_all_kernels.py > gpurun_out/r2_ncu_all_kernels_v5.log 2>&1)
python tools/ncu_kernel_table.py gpurun_out/r2_ncu_all_kernels_v5.csv --skip-first-half > gpurun_out/r2_ncu_all_kernels_v5.txt 2>&1
grep -E "relu_bwd|gru_out_bwd|gru_ru_bwd|^#" gpurun_out/r2_ncu_all_kernels_v5.txt
(RAMNET_COLSUM_ATOMIC=1 timeout 400 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null) > gpurun_out/r2c21_train_atomic.json
(timeout 400 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null) > gpurun_out/r2c21_train_lastblock.json
for f in atomic lastblock; do python -c "
import json
t=json.loads(open('gpurun_out/r2c21_train_$f.json').read().strip().splitlines()[-1])
print('$f train', round(t['maps_per_s'],1), 'ms', round(t['ms_per_step'],2), 'loss', t['loss'])"; done
