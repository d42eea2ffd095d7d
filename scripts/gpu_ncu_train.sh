#!/bin/bash
# ncu launch list of the training step alone (scripts/time_train.py: 2 warm-up steps + 1 timed), bounded
set -u
mkdir -p gpurun_out
timeout 200 python scripts/time_train.py 1 > gpurun_out/plain_train.log 2>&1 && \
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python scripts/time_train.py 1 > gpurun_out/ncu_train.log 2>&1
echo "ncu exit $?"; tail -c 300 gpurun_out/plain_train.log
