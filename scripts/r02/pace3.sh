#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/pace3.log
for at in 224 304; do
for dbg in 64 0; do
  echo "== atoms $at MDSCTK_TC_DEBUG=$dbg" >> gpurun_out/pace3.log
  ATOMS=$at MDSCTK_TC_DEBUG=$dbg MDSCTK_TC_PROF=1 MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so VERSIONS="2" ONLY=C3 timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "tc2 prof" | cut -c1-330 | tail -2 >> gpurun_out/pace3.log
done; done
cat gpurun_out/pace3.log
