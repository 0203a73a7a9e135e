#!/bin/bash
# clock-counter runs of the sweep (scripts/build_prof.sh first); logs into gpurun_out/prof.log
mkdir -p gpurun_out
LOG=gpurun_out/prof.log
: > $LOG
export MDSCTK_KNN_LIBRARY=$PWD/scripts/probe/libmdsctk_knn_prof.so MDSCTK_TC_PROF=1

ITER_REPS=1 ITER_DBG="${PROF_DBG:-6:0 6:8 6:1 4:0}" timeout 300 python scripts/gpu_iter.py >> $LOG 2>&1
if [ -n "$PROF_BASINS" ]; then
echo "=== basins $PROF_BASINS" >> $LOG
ITER_BASINS=$PROF_BASINS ITER_REPS=1 ITER_DBG="${PROF_DBG2:-6:0 6:1}" timeout 300 python scripts/gpu_iter.py >> $LOG 2>&1
fi
cat $LOG
