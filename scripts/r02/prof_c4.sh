#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/prof_c4.log
: > $LOG
export MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so MDSCTK_TC_PROF=1 VERSIONS="2" C4ROWS=131072
for spec in "C4 0" "C4 128" "C3 0" "C3 128"; do
  set -- $spec
  echo "== $1 MDSCTK_TC_DEBUG=$2" >> $LOG
  ONLY=$1 MDSCTK_TC_DEBUG=$2 timeout 300 python scripts/r02/time_sweep.py 2>&1 | tail -2 >> $LOG
done
cat $LOG
