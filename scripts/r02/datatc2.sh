#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_data_tc.py tests/test_gpu_parity.py tests/test_sparse.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/datatc.log
STREAMING=0 N=1000000 ONE_BLOCK=1 timeout 300 python scripts/r02/time_data.py 2>&1 | tail -2 >> gpurun_out/datatc.log
STREAMING=0 N=200000 timeout 300 python scripts/r02/time_data.py 2>&1 | tail -2 >> gpurun_out/datatc.log
MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so MDSCTK_TC_PROF=1 STREAMING=0 N=1000000 ONE_BLOCK=1 timeout 300 python scripts/r02/time_data.py 2>&1 | grep "data prof" | tail -2 >> gpurun_out/datatc.log
cat gpurun_out/datatc.log
