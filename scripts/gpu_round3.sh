#!/bin/bash
# round-end measurement at HEAD: tests, smoke, both bench arms, then the ncu launch list of the train step
set -u
bash scripts/gpu_round2.sh
bash scripts/gpu_ncu_train.sh
