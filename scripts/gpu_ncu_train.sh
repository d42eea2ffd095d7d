#!/bin/bash
# ncu launch list of the training step (small eval part, 1 timed train step)
set -u
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 3 --news 3000 --impressions 300 --no-cpu-baseline --train-steps 1"
timeout 600 python bench.py $ARGS > gpurun_out/plain_train.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python bench.py $ARGS > gpurun_out/ncu_train.log 2>&1
echo "ncu exit $?"; tail -c 600 gpurun_out/plain_train.log
