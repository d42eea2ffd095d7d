#!/bin/bash
# tensor-core encoder modes: kernel + encoder parity tests, then Stage A timing A/B (small layers by FFMA vs fp32x3 passes)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "mha or x3 or encoder or bf16 or mode or pair" > gpurun_out/pytest_x3.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_x3.log
timeout 300 python scripts/time_stage_a.py > gpurun_out/stage_a_x3.log 2>&1; echo "stage a exit $?"; cat gpurun_out/stage_a_x3.log | tail -9
