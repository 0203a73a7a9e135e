#!/bin/bash
# all-light passes (experiment bit 64) against the atom count: where does the cost per MMA jump?
mkdir -p gpurun_out
: > gpurun_out/atoms_scan.log
for at in 208 224 240 256 272 288 304; do
  echo "== atoms $at" >> gpurun_out/atoms_scan.log
  ATOMS=$at MDSCTK_TC_DEBUG=64 MDSCTK_TC_PROF=1 MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so VERSIONS="2" ONLY=C3 timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "tc2 prof" | cut -c1-330 | tail -1 >> gpurun_out/atoms_scan.log
done
cat gpurun_out/atoms_scan.log
