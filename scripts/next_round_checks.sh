#!/usr/bin/env bash
# First GPU call of the next round: validate the two paths that were written without GPU time at the end of round 1.
#   gpurun --timeout 900 -- 'bash scripts/next_round_checks.sh'
set -u
mkdir -p gpurun_out
echo "== frame-rate conditioning (csrc/pwg_fc.cu, PK_PWG_FRAME_COND=1) =="
timeout 300 python scripts/gpu_check_pwg_fc.py 2>&1 | tee gpurun_out/pwg_fc_check.log | tail -12
echo "== training step as a CUDA graph per batch shape (PK_TRAIN_GRAPH=1) =="
PK_TRAIN_GRAPH=1 timeout 300 python -m pytest tests/test_gpu_training.py -x -q -m gpu 2>&1 | tail -4
timeout 120 python scripts/bench_train.py --steps 10 2>&1 | tail -1 | cut -c1-160
PK_TRAIN_GRAPH=1 timeout 120 python scripts/bench_train.py --steps 10 2>&1 | tail -1 | cut -c1-160
