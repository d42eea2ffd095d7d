#!/bin/bash
# ncu --set full of the scoring kernel on the bench workload (diag_nodes.py drives it), plus the error diagnostic
set -u
mkdir -p gpurun_out
timeout 300 python scripts/diag_score_error.py > gpurun_out/diag_err.log 2>&1; echo "diag_err exit $?"; tail -4 gpurun_out/diag_err.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_score_tc2 python scripts/diag_nodes.py > gpurun_out/ncu_tc2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_tc2.log
