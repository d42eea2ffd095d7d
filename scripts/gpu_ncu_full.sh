#!/bin/bash
# full-size bench workload: ncu launch list + one --set full capture of the scoring kernel (traffic for the roofline)
set -u
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 3 --no-cpu-baseline --train-steps 0"
timeout 900 python bench.py $ARGS > gpurun_out/plain_full.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_score_tc_full python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
