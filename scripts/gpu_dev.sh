#!/bin/bash
# development round on the GPU box: scoring tests first (short timeout: a deadlocked kernel must not hold the box), then the
# timing diagnostic on the bench workload (phase clocks from the PHASE_CLOCKS variant if it was built), then a short bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "tensor_core_scoring or cached_impression or sweep_shapes or bucket_sweep or properties" > gpurun_out/dev_score.log 2>&1
echo "score tests exit $?"; tail -30 gpurun_out/dev_score.log
if [ -f lime_cikm25_b200/liblime_b200_clk.so ]; then
LIME_B200_LIB=$PWD/lime_cikm25_b200/liblime_b200_clk.so timeout 300 python scripts/diag_nodes.py > gpurun_out/diag_nodes.log 2>&1; echo "diag exit $?"; tail -8 gpurun_out/diag_nodes.log
fi
timeout 300 python scripts/prof_score.py > gpurun_out/prof_plain.log 2>&1; echo "prof_score exit $?"; tail -2 gpurun_out/prof_plain.log
if [ "${1:-}" = "bench" ]; then
timeout 600 python bench.py --train-steps 0 --no-cpu-baseline > gpurun_out/bench_dev.json 2> gpurun_out/bench_dev.err; echo "bench exit $?"; cat gpurun_out/bench_dev.json; tail -5 gpurun_out/bench_dev.err
fi
if [ "${1:-}" = "full" ]; then
timeout 900 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
fi
