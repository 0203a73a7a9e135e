#!/bin/bash
# build a variant of the library with extra nvcc flags -> scripts/probe/libmdsctk_knn_<name>.so
# use:  scripts/build_variant.sh q32 -DMDSCTK_TC2_QCAP=32 ;  MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_q32.so python ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p scripts/probe/var_$name
for f in mdsctk_b200/csrc/*.cu; do
  o=scripts/probe/var_$name/$(basename ${f%.cu}).o
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ccbin /usr/bin/g++ "$@" -c $f -o $o &
done
wait
nvcc -shared -o scripts/probe/libmdsctk_knn_$name.so scripts/probe/var_$name/*.o -ccbin /usr/bin/g++ -cudart static
ls -la scripts/probe/libmdsctk_knn_$name.so
