#!/bin/bash
# launch list (+ optional full capture of one kernel) of a small bench run:  scripts/gpu_ncu.sh [kernel_regex]
set -u
mkdir -p gpurun_out
SMALL="--steps 2 --warmup 3 --news 8000 --impressions 8000 --no-cpu-baseline --train-steps 0"
timeout 600 python bench.py $SMALL > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py $SMALL > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"; cat gpurun_out/plain.log | tail -2
if [ -n "${1:-}" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s 3 -c 1 -f -o gpurun_out/prof_$1 python bench.py $SMALL > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
