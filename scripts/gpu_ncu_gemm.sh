#!/bin/bash
# ncu --set full of the tcgen05 GEMMs inside one bf16 training step (tensor-pipe utilisation evidence)
set -u
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 3 --news 3000 --impressions 300 --no-cpu-baseline --train-steps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_general_kernel -s 60 -c 3 -f -o gpurun_out/prof_gemm_bf16_bwd python bench.py $ARGS > gpurun_out/ncu_gemm1.log 2>&1
echo "ncu bwd exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_bf16_kernel -s 30 -c 3 -f -o gpurun_out/prof_gemm_bf16_fwd python bench.py $ARGS > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu fwd exit $?"
