#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/datapf.log
for pf in 0 4 16 2064 2052; do
  echo "PF=$pf" >> gpurun_out/datapf.log
  MDSCTK_DATA_PF=$pf MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so MDSCTK_TC_PROF=1 STREAMING=0 N=1000000 ONE_BLOCK=1 timeout 300 python scripts/r02/time_data.py 2>&1 | grep "data prof\|metric 0" | tail -2 >> gpurun_out/datapf.log
done
cat gpurun_out/datapf.log
