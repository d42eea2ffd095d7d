#!/bin/bash
# development round on the GPU box: scoring tests first (short timeout: a deadlocked kernel must not hold the box), then the
# whole GPU suite, then the timing diagnostic on the bench workload
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "tensor_core_scoring or cached_impression or sweep_shapes or bucket_sweep or properties" > gpurun_out/dev_score.log 2>&1
echo "score tests exit $?"; tail -30 gpurun_out/dev_score.log
timeout 300 python scripts/diag_nodes.py > gpurun_out/diag_nodes.log 2>&1; echo "diag exit $?"; tail -8 gpurun_out/diag_nodes.log
if [ "${1:-}" = "full" ]; then
timeout 900 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
fi
