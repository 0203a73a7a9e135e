#!/bin/bash
# round-2 ncu evidence: launch list of the default bench + `--set full` captures of the dominant kernels, exported as
# raw CSV on the box (the .ncu-rep files stay in /tmp except the C3 sweep's, which carries source-level data)
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 > gpurun_out/bench_under_ncu_r02.log 2>&1
full() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  $NCU --set full --import-source on -k regex:$k -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > gpurun_out/ncu_r02_$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu_r02_$name.csv 2>/dev/null
}
ONLY=C3 VERSIONS=2 full sweep_c3 rms_sweep_tc2 1 python scripts/r02/time_sweep.py
cp /tmp/ncu/sweep_c3.ncu-rep gpurun_out/ncu_r02_sweep_c3.ncu-rep
ONLY=C4 VERSIONS=2 full sweep_c4 rms_sweep_tc2 1 python scripts/r02/time_sweep.py
ONLY=single VERSIONS=2 full sweep_single_basin rms_sweep_tc2 1 python scripts/r02/time_sweep.py
ONLY=C4 VERSIONS=2 full pack_c4 pack_frames 0 python scripts/r02/time_sweep.py
ONLY=C4 VERSIONS=2 full rms_rescore_c4 rms_rescore_kernel 1 python scripts/r02/time_sweep.py
N=1000000 ONE_BLOCK=1 full data_sweep_c5 data_sweep_tc 1 python scripts/r02/time_data.py
N=1000000 ONE_BLOCK=1 full data_rescore_c5 data_rescore 1 python scripts/r02/time_data.py
N=1000000 ONE_BLOCK=1 full data_pack_c5 data_pack 0 python scripts/r02/time_data.py
ls -la gpurun_out/ncu_r02_* gpurun_out/launches_r02.csv
tail -3 gpurun_out/bench_under_ncu_r02.log | cut -c1-600
