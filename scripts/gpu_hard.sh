#!/bin/bash
# the hardest rows of C4 for the certificate: basin 6 (G = 17.5 nm^2), reproduced at full density by a 250k-frame,
# 16-basin trajectory of the same seed (frames 93750..109375 are identical to C4's); k=64
mkdir -p gpurun_out
LOG=gpurun_out/hard.log
: > $LOG
for sl in ${SLACKS:-130}; do
  echo "=== slack $sl" >> $LOG
  ITER_SLACK=$sl ITER_N=250000 ITER_BASINS=16 ITER_SEED=20260118 ITER_K1=65 ITER_ROWS=15625 ITER_ROW0=93750 ITER_REPS=2 ITER_DBG="${HARD_DBG:-6:0 4:0}" timeout 400 python scripts/gpu_iter.py >> $LOG 2>&1
done
cat $LOG
