#!/bin/bash
# timing of every experiment build lime_cikm25_b200/liblime_b200_<variant>.so on the bench workload (scripts/prof_score.py)
set -u
mkdir -p gpurun_out
for f in lime_cikm25_b200/liblime_b200*.so; do
  echo "== $f"; LIME_B200_LIB=$PWD/$f timeout 200 python scripts/prof_score.py 2>&1 | tail -1
done 2>&1 | tee gpurun_out/variants.log
