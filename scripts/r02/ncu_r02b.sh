#!/bin/bash
# round-2 ncu evidence, refreshed after the wide-stage sweep and the resident knn_data filter (same recipe as ncu_r02.sh,
# the three dominant kernels only): launch list of the default bench + `--set full` captures, raw CSV exported on the box
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 1 --warmup 1 > gpurun_out/bench_under_ncu_r02b.log 2>&1
full() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  $NCU --set full --import-source on -k regex:$k -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/ncu_r02b_$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu_r02b_$name.csv 2>/dev/null
}
ONLY=C3 VERSIONS=2 full sweep_c3 rms_sweep_tc2 1 python scripts/r02/time_sweep.py
ONLY=C4 VERSIONS=2 full sweep_c4 rms_sweep_tc2 1 python scripts/r02/time_sweep.py
N=1000000 ONE_BLOCK=1 full data_sweep_c5 data_sweep_tc 1 python scripts/r02/time_data.py
ls -la gpurun_out/ncu_r02b_* gpurun_out/launches_r02b.csv
tail -3 gpurun_out/bench_under_ncu_r02b.log | cut -c1-300
