#!/bin/bash
# round-end style run: the whole GPU suite, smoke(), then the bench with the driver's flags
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -12 > gpurun_out/full_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 >> gpurun_out/full_tests.log
cat gpurun_out/full_tests.log
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps ${STEPS:-20} --warmup ${WARMUP:-5} ) > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err
tail -4 gpurun_out/bench_ref_r02.err
( time timeout 2400 python bench.py --gpus 1 --steps ${STEPS:-20} --warmup ${WARMUP:-5} ) > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
tail -4 gpurun_out/bench_r02_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_ref_r02.json", "gpurun_out/bench_r02_n1.json"):
    for ln in open(f):
        if ln.startswith("{"):
            r = json.loads(ln)
            print(f, {k: r.get(k) for k in ("value", "ms_per_step", "n_gpus", "lists_hash")}, "e2e", r["e2e"]["value"], "roofline", r.get("roofline", {}).get("frac"), "parity ok", (r.get("parity") or {}).get("ok"))
            for s in r.get("secondary", []):
                print("   secondary", s["config"]["workload"][:50], "%.3e" % s["value"], "frac", round(s["roofline"]["frac"], 3), "parity", (s.get("parity") or {}).get("ok"), s.get("lists_hash"))
PY
