#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 4 --steps 10 --warmup 3 2> gpurun_out/r2c57_bench4.err) > gpurun_out/r2c57_bench4.json
echo "rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2c57_bench4.json').read().strip().splitlines()[-1]); t=d['train']
print('fwd', d['n_gpus'], round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity']['max_rel_err'], 'clocks', d['clocks'])
print('train', round(t['maps_per_s'],1), 'ms', round(t['ms_per_step'],2), 'loss', t['loss'], 'allreduce', t['allreduce'])"
