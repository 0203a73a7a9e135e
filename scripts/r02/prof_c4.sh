#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/prof_c4.log
: > $LOG
export MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so MDSCTK_TC_PROF=1 VERSIONS="2" ONLY=C4
for dbg in 0; do
  echo "== C4 block MDSCTK_TC_DEBUG=$dbg" >> $LOG
  MDSCTK_TC_DEBUG=$dbg timeout 300 python scripts/r02/time_sweep.py 2>&1 | tail -2 >> $LOG
done
cat $LOG
unset MDSCTK_KNN_LIBRARY MDSCTK_TC_PROF
VERSIONS="2 2" ONLY=C timeout 600 python scripts/r02/time_sweep.py 2>&1 | tail -4
