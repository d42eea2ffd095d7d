#!/bin/bash
# training-path development loop on the GPU box: the training tests, then the step time with the TMA GEMMs on / off
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/train_tests.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/train_tests.log
timeout 300 python scripts/time_train.py 10 2>&1 | tail -2
LIME_TRAIN_TMA=0 timeout 300 python scripts/time_train.py 10 2>&1 | tail -2
