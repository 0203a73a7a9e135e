#!/bin/bash
# ring depth vs refine-queue size: QCAP=16 (6 ring stages at 304 atoms, production) against QCAP=32 (5 stages)
mkdir -p gpurun_out
: > gpurun_out/qcap.log
for lib in "" scripts/probe/libmdsctk_knn_q32.so; do
  echo "== library: ${lib:-production}" >> gpurun_out/qcap.log
  MDSCTK_KNN_LIBRARY=$lib VERSIONS="2 2" timeout 600 python scripts/r02/time_sweep.py 2>&1 | grep -v "^$" | tail -6 >> gpurun_out/qcap.log
done
cat gpurun_out/qcap.log
