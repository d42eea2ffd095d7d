#!/bin/bash
# Runs on the GPU box under gpurun: tests, smoke, bench (both arms), then the two ncu passes.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cat gpurun_out/bench_ref.json
if [ "${1:-}" = "ncu" ]; then
  SMALL="--steps 2 --warmup 3 --news 8000 --impressions 8000 --no-cpu-baseline --train-steps 0"
  timeout 600 python bench.py $SMALL > gpurun_out/plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py $SMALL > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  timeout 600 python bench.py $SMALL > gpurun_out/plain2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 3 -c 2 -f -o gpurun_out/score_prof python bench.py $SMALL > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
