#!/bin/bash
# what paces a stage at ~800 clk?  all-light passes; bits: 64 all light, 128 stage pairs, 4096 spinning producer
mkdir -p gpurun_out
: > gpurun_out/pace.log
for at in 224 304; do
for dbg in 64 4160 192 4288; do
  echo "== atoms $at MDSCTK_TC_DEBUG=$dbg" >> gpurun_out/pace.log
  ATOMS=$at MDSCTK_TC_DEBUG=$dbg MDSCTK_TC_PROF=1 MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so VERSIONS="2" ONLY=C3 timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "tc2 prof" | cut -c1-330 | tail -1 >> gpurun_out/pace.log
done; done
cat gpurun_out/pace.log
