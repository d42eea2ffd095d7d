#!/bin/bash
# ncu --set full of the scoring kernel on a bench-shaped workload (scripts/prof_score.py drives it): plain run first, then ncu
set -u
mkdir -p gpurun_out
N=${1:-30000}
timeout 300 python scripts/prof_score.py 65238 $N 3 > gpurun_out/prof_plain.log 2>&1; echo "plain exit $?"; tail -1 gpurun_out/prof_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_tc python scripts/prof_score.py 65238 $N 3 > gpurun_out/ncu_tc.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_tc.log
