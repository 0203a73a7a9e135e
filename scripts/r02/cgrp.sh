#!/bin/bash
# ring slots released per commit: 1 (old), 2, 3, 4 -- production-like timing (prof library without MDSCTK_TC_PROF) and counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/cgrp.log
for at in 304 224; do
for g in 1 2 3 4; do
  echo "== atoms $at commit group $g" >> gpurun_out/cgrp.log
  ATOMS=$at MDSCTK_TC_CGRP=$g MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so VERSIONS="2 2" ONLY=${ONLY:-C} timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "version" | awk 'NR%2==0' | cut -c1-120 >> gpurun_out/cgrp.log
  ATOMS=$at MDSCTK_TC_CGRP=$g MDSCTK_TC_DEBUG=64 MDSCTK_TC_PROF=1 MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so VERSIONS="2" ONLY=C3 timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "tc2 prof" | cut -c1-250 | tail -1 >> gpurun_out/cgrp.log
done; done
cat gpurun_out/cgrp.log
