#!/bin/bash
# compute-sanitizer (memcheck, then racecheck + synccheck) over the scoring tests; three plain repeats of the GPU suite for flakiness
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "dedup or tensor_core_scoring" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -3
done
for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -1; done
