#!/bin/bash
# full GPU suite + default bench (1 GPU); logs into gpurun_out/full.log
mkdir -p gpurun_out
LOG=gpurun_out/full.log
: > $LOG
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) >> $LOG 2>&1
python -c "import __graft_entry__ as g; g.smoke()" >> $LOG 2>&1
timeout 600 python bench.py > gpurun_out/bench_default.json 2>> $LOG
cat gpurun_out/bench_default.json >> $LOG
cat $LOG
