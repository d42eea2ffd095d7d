#!/bin/bash
# Stage A evidence: GEMM micro-bench, launch list of a bf16 cache build (16,384 news), one full ncu capture of gemm_tma_kernel
set -u
mkdir -p gpurun_out
timeout 200 python scripts/bench_gemm_tma.py > gpurun_out/gemm_tma.log 2>&1; echo "gemm bench exit $?"; tail -5 gpurun_out/gemm_tma.log
timeout 300 python scripts/time_stage_a.py 16384 > gpurun_out/stage_a_16k.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_a.csv python scripts/time_stage_a.py 16384 > gpurun_out/ncu_a.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 8 -c 1 -f -o gpurun_out/prof_gemm_tma python scripts/bench_gemm_tma.py > gpurun_out/ncu_gemm_tma.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_gemm_tma.log
