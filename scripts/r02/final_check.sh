#!/bin/bash
# last GPU call of the round: the tensor-path tests (every layout family of the sweep), the hardening suite, C3 / C4-block timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_hardening.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/final_check.log
VERSIONS="2 2" ONLY=C timeout 300 python scripts/r02/time_sweep.py 2>&1 | grep version | cut -c1-140 >> gpurun_out/final_check.log
ATOMS=224 VERSIONS="2 -2" ONLY=C3 timeout 200 python scripts/r02/time_sweep.py 2>&1 | grep version | cut -c1-140 >> gpurun_out/final_check.log
cat gpurun_out/final_check.log
