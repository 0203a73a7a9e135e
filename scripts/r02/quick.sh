#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_hardening.py -x -q 2>&1 | tail -3 > gpurun_out/quick.log
VERSIONS="${VERSIONS:-1 2 2}" ONLY="${ONLY:-}" timeout 600 python scripts/r02/time_sweep.py >> gpurun_out/quick.log 2>&1
timeout 1500 python bench.py --steps 2 --warmup 1 --secondary none > gpurun_out/bench_quick.json 2>> gpurun_out/quick.log
python - >> gpurun_out/quick.log <<'PY'
import json
r = json.loads([l for l in open("gpurun_out/bench_quick.json") if l.startswith("{")][0])
print({k: r[k] for k in ("value", "ms_per_step", "lists_hash")}, "e2e", r["e2e"]["value"], {k: round(r["roofline"][k], 3) for k in ("frac", "sweep_ms_per_step", "post_ms_per_step", "step_frac")}, r["parity"]["ok"], r["config"]["k_keep"])
PY
cat gpurun_out/quick.log
