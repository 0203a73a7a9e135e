#!/bin/bash
# round-2 validation on a GPU box: GPU test suite, then the default bench (N=1)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gputests.log
cat gpurun_out/gputests.log
timeout 1200 python bench.py ${BENCH_ARGS:---steps 2 --warmup 1} > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
