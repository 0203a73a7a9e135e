#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/segs.log
for sg in 4 8 12 16 24; do
  echo "== segments $sg" >> gpurun_out/segs.log
  C4ROWS=131072 SEGMENTS=$sg VERSIONS="2 2" ONLY=C4 timeout 300 python scripts/r02/time_sweep.py 2>&1 | tail -1 >> gpurun_out/segs.log
done
for sg in 4 6 8; do
  echo "== C3 segments $sg" >> gpurun_out/segs.log
  SEGMENTS=$sg VERSIONS="2 2" ONLY=C3 timeout 300 python scripts/r02/time_sweep.py 2>&1 | tail -1 >> gpurun_out/segs.log
done
cat gpurun_out/segs.log
