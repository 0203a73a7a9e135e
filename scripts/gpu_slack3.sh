#!/bin/bash
mkdir -p gpurun_out; LOG=gpurun_out/slack3.log; : > $LOG
for sl in 96 64 48 32; do
  echo "=== C3 slack $sl" >> $LOG
  ITER_SLACK=$sl ITER_REPS=2 ITER_DBG="6:0" timeout 300 python scripts/gpu_iter.py >> $LOG 2>&1
done
cat $LOG
