#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_data_tc.py tests/test_gpu_parity.py -x -q 2>&1 | tail -30 > gpurun_out/datatc.log
timeout 300 python scripts/r02/time_data.py >> gpurun_out/datatc.log 2>&1
cat gpurun_out/datatc.log
