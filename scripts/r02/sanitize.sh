#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --launch-timeout 900 --print-limit 20 python scripts/r02/sanitize.py > gpurun_out/sanitize.log 2>&1
tail -40 gpurun_out/sanitize.log
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k resident 2>&1 | tail -3
