#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/prof_rtt.log
: > $LOG
export MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so MDSCTK_TC_PROF=1 VERSIONS="2" ONLY=C3
for spec in "300 66" "300 64" "224 66"; do
  set -- $spec
  echo "== ATOMS=$1 MDSCTK_TC_DEBUG=$2" >> $LOG
  ATOMS=$1 MDSCTK_TC_DEBUG=$2 timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "prof\]" | tail -1 >> $LOG
done
cat $LOG
