#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r2c15_bench2.err) > gpurun_out/r2c15_bench2.json
echo "rc=$?"
tail -5 gpurun_out/r2c15_bench2.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r2c15_bench2.json').read().strip().splitlines()[-1]); t=d['train']
print('fwd', d['n_gpus'], round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'parity', d['parity'])
print('train', round(t['maps_per_s'],1), 'ms', round(t['ms_per_step'],2), 'loss', t['loss'], 'allreduce', t['allreduce'])"
(timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q 2>&1 | tail -3)
