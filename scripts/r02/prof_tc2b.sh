#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/prof_tc2b.log
: > $LOG
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -5 >> $LOG
echo "== production build" >> $LOG
VERSIONS="1 2 2" ONLY=${ONLY:-C3} timeout 300 python scripts/r02/time_sweep.py 2>&1 | tail -3 >> $LOG
export MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so MDSCTK_TC_PROF=1 VERSIONS="2" ONLY=C3
for spec in "300 0" "300 64" "300 66" "300 320" "300 322"; do
  set -- $spec
  echo "== ATOMS=$1 MDSCTK_TC_DEBUG=$2" >> $LOG
  ATOMS=$1 MDSCTK_TC_DEBUG=$2 timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "prof\]" | tail -1 >> $LOG
done
cat $LOG
