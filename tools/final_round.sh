#!/bin/bash
# Round-end measurement batch (run under gpurun): tests, bench lines, layer tables, latency, smoke.
set -x
mkdir -p gpurun_out
(timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/f_tests.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/f_clocks.csv &
SMI=$!
python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench_fwd.json 2> gpurun_out/f_bench_fwd.err
python bench.py --mode train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_train.json 2> gpurun_out/f_bench_train.err
kill $SMI
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
timeout 200 python tools/layer_bench.py > gpurun_out/f_layers.txt 2>&1
timeout 200 python tools/wgrad_bench.py > gpurun_out/f_wgrad.txt 2>&1
timeout 200 python tools/train_profile.py > gpurun_out/f_train_profile.txt 2>&1
timeout 200 python tools/latency_bench.py --json gpurun_out/f_latency.json > gpurun_out/f_latency.txt 2>&1
timeout 100 python tools/head_bench.py > gpurun_out/f_head.txt 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.txt 2>&1
tail -3 gpurun_out/f_tests.log gpurun_out/f_smoke.txt
