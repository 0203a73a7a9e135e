#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/tc2.log
: > $LOG
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -5 >> $LOG
timeout 900 python -m pytest tests/test_gpu_hardening.py tests/test_gpu_parity.py -x -q 2>&1 | tail -8 >> $LOG
timeout 300 python scripts/r02/time_sweep.py >> $LOG 2>&1
cat $LOG
