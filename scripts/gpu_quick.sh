#!/bin/bash
# quick GPU check of a subset of tests: scripts/gpu_quick.sh "<pytest -k expression>"
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -k "$1" > gpurun_out/quick.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/quick.log
