#!/bin/bash
# clean timing (no clock counters compiled in): experiment library, C3 and the C4 block, ms only.
# Record of the run quoted in profiles/README.md; MDSCTK_TC_CGRP (ring slots per commit) and the exp1 library (lane-0 polls,
# -DMDSCTK_TC2_POLL1=1) were experiment code of that working tree and are gone -- both measured neutral or worse.
mkdir -p gpurun_out
: > gpurun_out/clean.log
run() { echo "== $1 atoms=$2 dbg=$3 cgrp=$4" >> gpurun_out/clean.log
  ATOMS=$2 MDSCTK_TC_DEBUG=$3 MDSCTK_TC_CGRP=$4 MDSCTK_KNN_LIBRARY=scripts/probe/libmdsctk_knn_$1.so VERSIONS="2 2" ONLY=${5:-C3} timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep "version" | awk 'NR%2==0' | cut -c1-60 >> gpurun_out/clean.log; }
run exp 304 0 1 C
run exp 304 0 2 C
run exp1 304 0 1 C
run exp 304 64 1
run exp 304 64 2
run exp1 304 64 1
run exp 304 66 1
run exp 304 192 1
run exp 224 64 1
run exp 224 64 2
run exp 224 64 4
run exp1 224 64 1
run exp 224 66 1
run exp 224 192 1
cat gpurun_out/clean.log
