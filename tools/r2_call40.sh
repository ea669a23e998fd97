#!/bin/bash
# live norm layers: kernel tests, model goldens (train-mode BN / IN), gradient goldens
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "norm" > gpurun_out/r2c40_norm_kernels.log 2>&1; echo "norm kernels rc=$?"; tail -15 gpurun_out/r2c40_norm_kernels.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "bn_ or in_ or unet_bn" > gpurun_out/r2c40_norm_models.log 2>&1; echo "norm models rc=$?"; tail -25 gpurun_out/r2c40_norm_models.log
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -k "test_model_gradients_match_reference" > gpurun_out/r2c40_norm_grads.log 2>&1; echo "norm grads rc=$?"; tail -25 gpurun_out/r2c40_norm_grads.log
