#!/bin/bash
# bits: 64 all light, 2 no MMAs, 1024 no scout (all tiles live), 128 stage pairs
mkdir -p gpurun_out
: > gpurun_out/pace2.log
for at in 224; do
for dbg in 66 1090 194 1088; do
  echo "== atoms $at MDSCTK_TC_DEBUG=$dbg" >> gpurun_out/pace2.log
  ATOMS=$at MDSCTK_TC_DEBUG=$dbg MDSCTK_TC_PROF=1 MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so VERSIONS="2" ONLY=C3 timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "tc2 prof" | cut -c1-330 | tail -1 >> gpurun_out/pace2.log
done; done
cat gpurun_out/pace2.log
