#!/bin/bash
# ncu launch lists of one cache build per tensor-core encoder mode (16,384 news = 2 encoder passes, built twice: the second is warm)
set -u
mkdir -p gpurun_out
for mode in bf16 x3; do
  timeout 120 python scripts/time_stage_a.py 16384 0 $mode > gpurun_out/plain_$mode.log 2>&1 && \
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$mode.csv python scripts/time_stage_a.py 16384 0 $mode > gpurun_out/ncu_$mode.log 2>&1
  echo "$mode ncu exit $?"; tail -2 gpurun_out/plain_$mode.log
done
