#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/explore1.log
: > $LOG
nvidia-smi --query-gpu=name,memory.total --format=csv >> $LOG
echo "== C3 floors (16 basins)" >> $LOG
ITER_N=100000 ITER_DBG="6:0 6:1 6:5 6:3 6:7 6:1025 6:0" ITER_REPS=1 timeout 300 python scripts/gpu_iter.py >> $LOG 2>&1
echo "== C3 single basin" >> $LOG
ITER_N=100000 ITER_BASINS=1 ITER_DBG="6:0 6:1 6:5" ITER_REPS=1 timeout 300 python scripts/gpu_iter.py >> $LOG 2>&1
echo "== C5" >> $LOG
timeout 600 python scripts/r02/explore1.py c5 >> $LOG 2>&1
echo "== C4" >> $LOG
timeout 900 python scripts/r02/explore1.py c4 >> $LOG 2>&1
tail -60 $LOG
