#!/bin/bash
# effect of the candidate slack on the C4-shaped sweep (200k frames, 64 basins, k=64, 16384 rows from ITER_ROW0)
mkdir -p gpurun_out
LOG=gpurun_out/slack.log
: > $LOG
for sl in 130 195 260; do
  echo "=== slack $sl" >> $LOG
  ITER_SLACK=$sl ITER_N=200000 ITER_BASINS=64 ITER_SEED=20260118 ITER_K1=65 ITER_ROWS=16384 ITER_ROW0=${ROW0:-16384} ITER_REPS=2 ITER_DBG="6:0" timeout 300 python scripts/gpu_iter.py >> $LOG 2>&1
done
cat $LOG
