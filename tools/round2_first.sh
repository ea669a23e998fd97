#!/bin/bash
# First GPU call of round 2: validate and A/B the paths that were written after round 1's GPU budget was spent
# (all default OFF).  Run under gpurun; writes gpurun_out/r2_*.  Every step is bounded by `timeout`.
mkdir -p gpurun_out
export RAMNET_TEST_EXPERIMENTAL=1
(RAMNET_WGRAD_FOLD=1 RAMNET_WGRAD_FUSED_SUM=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -m gpu -q 2>&1 | tail -30) > gpurun_out/r2_experimental_tests.log
# weight gradient: folding / fused split sum, per layer
(timeout 120 python tools/wgrad_bench.py 2>&1) > gpurun_out/r2_wgrad_base.txt
(RAMNET_WGRAD_FOLD=1 timeout 120 python tools/wgrad_bench.py --only dec2 2>&1) > gpurun_out/r2_wgrad_fold.txt
(RAMNET_WGRAD_FUSED_SUM=1 timeout 120 python tools/wgrad_bench.py 2>&1) > gpurun_out/r2_wgrad_fused_sum.txt
# forward: hpack for the Cout = 32 layers (dec2) -- layer_bench packs through ops.pack_weights, so time it through the model
(timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null) > gpurun_out/r2_bench_base.json
(RAMNET_HPACK=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2_bench_hpack.err) > gpurun_out/r2_bench_hpack.json
(RAMNET_HPACK=1 timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2_model_tests_hpack.log
(RAMNET_WGRAD_FOLD=1 RAMNET_WGRAD_FUSED_SUM=1 timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2_train_tests_fold.log
(RAMNET_WGRAD_FOLD=1 RAMNET_WGRAD_FUSED_SUM=1 RAMNET_HPACK=1 timeout 300 python bench.py --mode train --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null) > gpurun_out/r2_bench_train_all.json
tail -3 gpurun_out/r2_experimental_tests.log gpurun_out/r2_model_tests_hpack.log gpurun_out/r2_train_tests_fold.log 2>/dev/null | cat
# sanitizer feasibility probe (memcheck on the smoke path)
(timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke 2>&1 | tail -40) > gpurun_out/r2_sanitizer_memcheck_smoke.log
tail -5 gpurun_out/r2_sanitizer_memcheck_smoke.log
