#!/bin/bash
# clock-counter build of the library (MDSCTK_TC_PROF_BUILD=1) -> scripts/probe/libmdsctk_knn_prof.so
# use:  MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so MDSCTK_TC_PROF=1 python scripts/gpu_iter.py
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/probe/prof_obj
for f in mdsctk_b200/csrc/*.cu; do
  o=scripts/probe/prof_obj/$(basename ${f%.cu}).o
  if [ "$f" -nt "$o" ] || [ mdsctk_b200/csrc/common.cuh -nt "$o" ] || [ "$(basename $f)" = rms_tc.cu ]; then
    nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ccbin /usr/bin/g++ \
         -DMDSCTK_TC_PROF_BUILD=1 -DMDSCTK_TC_EXPERIMENTS=1 -c $f -o $o &
  fi
done
wait
nvcc -shared -o scripts/probe/libmdsctk_knn_prof.so scripts/probe/prof_obj/*.o -ccbin /usr/bin/g++ -cudart static
ls -la scripts/probe/libmdsctk_knn_prof.so
