#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/prof_n.log
: > $LOG
export MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_prof.so MDSCTK_TC_PROF=1 VERSIONS="2" ONLY=C3
for n in 32 48 64 80 96 112 128 144 160 176 192 208 224 240 256; do
  echo "== N=$n" >> $LOG
  MDSCTK_TC_N=$n ATOMS=128 MDSCTK_TC_DEBUG=64 timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "prof\]" | tail -1 >> $LOG
done
cat $LOG
