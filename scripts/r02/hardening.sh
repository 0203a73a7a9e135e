#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_hardening.py -x -q -s 2>&1 | tail -40 > gpurun_out/hardening.log
cat gpurun_out/hardening.log
