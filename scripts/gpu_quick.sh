#!/bin/bash
# quick GPU check: tensor-core parity tests + C3 timings of the settings in QUICK_DBG; logs into gpurun_out/quick.log
mkdir -p gpurun_out
LOG=gpurun_out/quick.log
: > $LOG
if [ "${QUICK_TESTS:-1}" = "1" ]; then
( timeout 420 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -5 ) >> $LOG 2>&1
fi
ITER_REPS=2 ITER_DBG="${QUICK_DBG:-6:0 6:256}" timeout 300 python scripts/gpu_iter.py >> $LOG 2>&1
if [ -n "$QUICK_C4" ]; then
ITER_N=200000 ITER_BASINS=64 ITER_SEED=20260118 ITER_K1=65 ITER_ROWS=16384 ITER_REPS=1 ITER_DBG="$QUICK_C4" timeout 300 python scripts/gpu_iter.py >> $LOG 2>&1
fi
cat $LOG
