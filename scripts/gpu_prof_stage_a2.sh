#!/bin/bash
# ncu launch list of the cache build in one tensor-core encoder mode (16,384 news = 2 encoder passes, built twice: the second is warm)
set -u
mkdir -p gpurun_out
for mode in ${1:-bf16 x3}; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$mode.csv python scripts/time_stage_a.py 16384 0 $mode > gpurun_out/ncu_$mode.log 2>&1
  echo "$mode ncu exit $?"
done
