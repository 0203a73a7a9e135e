#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/quick.log
VERSIONS="${VERSIONS:--2 2 -2 2}" ONLY="${ONLY:-C}" timeout 600 python scripts/r02/time_sweep.py >> gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
