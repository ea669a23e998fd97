#!/bin/bash
# 2-GPU validation of the driver's command after the two-stream changes (front stream, weight-gradient stream) + single-stream roofline instrumentation
mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r2c43_bench2.err) > gpurun_out/r2c43_bench2.json
echo "rc=$?"; tail -3 gpurun_out/r2c43_bench2.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r2c43_bench2.json').read().strip().splitlines()[-1]); t=d['train']
print('fwd', d['n_gpus'], round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity']['max_rel_err'], 'frac', d['roofline']['frac'], d['roofline']['frac_of_tf32_pipe'], d['roofline']['frac_of_step'])
print('train', round(t['maps_per_s'],1), 'ms', round(t['ms_per_step'],2), 'loss', t['loss'], 'allreduce', t['allreduce'])
print({k:t[k] for k in ('conv_fwd_dgrad','wgrad','other_kernels_ms_per_step')})"
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2> gpurun_out/r2c43_ref2.err) > gpurun_out/r2c43_ref2.json; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2c43_ref2.json
