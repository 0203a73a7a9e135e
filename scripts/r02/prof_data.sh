#!/bin/bash
# clock-counter profile of the knn_data filter (prof build): one 131072 x 1M x 512 block
mkdir -p gpurun_out
MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so MDSCTK_TC_PROF=1 STREAMING=0 N=1000000 ONE_BLOCK=1 timeout 300 python scripts/r02/time_data.py 2>&1 | grep -v "^$" | tail -8 > gpurun_out/prof_data.log
cat gpurun_out/prof_data.log
