#!/bin/bash
# all-light passes (experiment bit 64): ring depth 5 (QCAP=32) against 6 (QCAP=16), clock counters
mkdir -p gpurun_out
: > gpurun_out/qcap2.log
for lib in scripts/probe/libmdsctk_knn_prof.so scripts/probe/libmdsctk_knn_q16x.so; do
  for dbg in 64 0; do
    echo "== $lib MDSCTK_TC_DEBUG=$dbg" >> gpurun_out/qcap2.log
    MDSCTK_TC_DEBUG=$dbg MDSCTK_TC_PROF=1 MDSCTK_KNN_LIBRARY=$lib VERSIONS="2" ONLY=C timeout 600 python scripts/r02/time_sweep.py 2>&1 | grep "tc2 prof\|version" | cut -c1-330 | awk 'NR%2==0 || /version/' | tail -6 >> gpurun_out/qcap2.log
  done
done
cat gpurun_out/qcap2.log
