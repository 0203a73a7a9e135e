#!/bin/bash
# round-1d GPU session: parity of the 1xFP16 sweep, then timings (logs into gpurun_out/)
mkdir -p gpurun_out
LOG=gpurun_out/r01d.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader >> $LOG 2>&1
( time timeout 420 python -m pytest tests/test_gpu_tc.py -x -q -s 2>&1 | grep -E "tc kernel|passed|failed|Error|error|assert" | tail -40 ) >> $LOG 2>&1
echo "=== C3 iter" >> $LOG
ITER_REPS=2 ITER_DBG="4:0 6:0 5:0 6:1 6:3 6:5" timeout 300 python scripts/gpu_iter.py >> $LOG 2>&1
echo "=== C4-shape (200k frames, 64 basins, k=64, 16384 rows)" >> $LOG
ITER_N=200000 ITER_BASINS=64 ITER_SEED=20260118 ITER_K1=65 ITER_ROWS=16384 ITER_REPS=1 ITER_DBG="4:0 6:0" timeout 300 python scripts/gpu_iter.py >> $LOG 2>&1
echo "=== bench kernel 6" >> $LOG
BENCH_KNN_DATA=0 timeout 300 python bench.py --rms-kernel 6 > gpurun_out/bench_k6.json 2>> $LOG
cat gpurun_out/bench_k6.json >> $LOG
cat $LOG
