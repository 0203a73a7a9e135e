#!/bin/bash
# experiment: extra tcgen05.commit per round (bit 16384: +1, 32768: +2) in all-light mode
mkdir -p gpurun_out
: > gpurun_out/pace5.log
for at in 224; do
for dbg in 64 16448 32832 49216; do
  echo "== atoms $at MDSCTK_TC_DEBUG=$dbg" >> gpurun_out/pace5.log
  ATOMS=$at MDSCTK_TC_DEBUG=$dbg MDSCTK_TC_PROF=1 MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so VERSIONS="2" ONLY=C3 timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "tc2 prof" | cut -c1-420 | tail -2 >> gpurun_out/pace5.log
done; done
cat gpurun_out/pace5.log
