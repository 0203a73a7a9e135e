#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/segs_data.log
for sg in 1 2 3 4 6 8; do
  echo "== data segments $sg" >> gpurun_out/segs_data.log
  SEGMENTS=$sg N=1000000 ONE_BLOCK=1 timeout 300 python scripts/r02/time_data.py 2>&1 | tail -1 >> gpurun_out/segs_data.log
done
cat gpurun_out/segs_data.log
