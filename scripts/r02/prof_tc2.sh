#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/prof_tc2.log
: > $LOG
export MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so MDSCTK_TC_PROF=1
for dbg in ${DBGS:-0 2 1024}; do
  echo "== MDSCTK_TC_DEBUG=$dbg" >> $LOG
  MDSCTK_TC_DEBUG=$dbg ONLY=${ONLY:-C3} timeout 300 python scripts/r02/time_sweep.py >> $LOG 2>&1
done
cat $LOG
