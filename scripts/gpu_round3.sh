#!/bin/bash
# round-end measurement at HEAD: tests, smoke, both bench arms, then ncu launch lists of the bench and of the train step
set -u
bash scripts/gpu_round2.sh
ARGS="--steps 2 --warmup 3 --no-cpu-baseline --train-steps 0"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
bash scripts/gpu_ncu_train.sh
