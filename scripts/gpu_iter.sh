#!/bin/bash
# iteration helper: runs on the GPU box; logs into gpurun_out/iter.log
mkdir -p gpurun_out
LOG=gpurun_out/iter.log
: > $LOG
if [ "${ITER_TESTS:-1}" = "1" ]; then
  timeout 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -25 >> $LOG
fi
timeout ${ITER_TIMEOUT:-240} python scripts/gpu_iter.py >> $LOG 2>&1
cat $LOG
