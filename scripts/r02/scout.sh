#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/scout.log
VERSIONS="2 2" timeout 600 python scripts/r02/time_sweep.py 2>&1 | grep version >> gpurun_out/scout.log
for at in 224 304; do
for dbg in 64 0; do
  echo "== atoms $at MDSCTK_TC_DEBUG=$dbg" >> gpurun_out/scout.log
  ATOMS=$at MDSCTK_TC_DEBUG=$dbg MDSCTK_TC_PROF=1 MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so VERSIONS="2" ONLY=C3 timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "tc2 prof" | cut -c1-330 | tail -2 >> gpurun_out/scout.log
done; done
cat gpurun_out/scout.log
