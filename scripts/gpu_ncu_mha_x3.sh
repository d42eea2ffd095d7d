#!/bin/bash
# one ncu --set full capture of the fp32x3 attention kernels (T = 32 and T = 128) inside a small cache build
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mha_x3 -c 2 -f -o gpurun_out/prof_mha_x3 python scripts/time_stage_a.py 8192 0 x3 > gpurun_out/ncu_mha_x3.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_mha_x3.log
