#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/wide.log
run() { echo "== $1 atoms=$2 dbg=$3 wide=$4" >> gpurun_out/wide.log
  ATOMS=$2 MDSCTK_TC_DEBUG=$3 MDSCTK_TC_WIDE=$4 MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_$1.so VERSIONS="-2 -2" ONLY=${5:-C3} timeout 400 python scripts/r02/time_sweep.py 2>&1 | grep "version" | awk 'NR%2==0' | cut -c1-75 >> gpurun_out/wide.log; }
run exp 304 0 1 C
run exp 304 0 0 C
run exp 304 64 1
run exp 304 64 0
run exp 224 0 1
run exp 224 0 0
run exp 224 64 1
run exp 224 64 0
cat gpurun_out/wide.log
