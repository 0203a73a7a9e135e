#!/bin/bash
# eight-GPU validation: the C++ tools' --gpus path at 2 and 4 GPUs, bench.py under torchrun at N=8 and N=4
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/n8.log
timeout 900 python -m pytest tests/test_tools.py -x -q -m gpu -k "several_gpus" 2>&1 | tail -5 >> gpurun_out/n8.log
for n in 8 4; do
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 1 --secondary c5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
tail -c 600 gpurun_out/bench_n$n.err >> gpurun_out/n8.log
done
cat gpurun_out/n8.log
python - <<'PY'
import json
for n in (8, 4):
  for ln in open(f"gpurun_out/bench_n{n}.json"):
    if ln.startswith("{"):
        r = json.loads(ln)
        print(n, {k: r[k] for k in ("value", "ms_per_step", "n_gpus", "lists_hash")}, "parity", r["parity"]["ok"], "e2e", r["e2e"]["value"],
              "roofline", {k: round(r["roofline"][k], 3) for k in ("achieved", "frac", "sweep_ms_per_step", "step_frac")},
              {k: r["config"][k] for k in ("allgather_ms", "allgather_bytes_per_gpu", "h2d_ms", "pack_ms", "fallback_rows", "fit_rows_per_rank_per_step")})
        for s in r["secondary"]:
            print("   secondary", s["config"]["workload"][:30], "%.3e" % s["value"], s["lists_hash"], s["parity"]["ok"], "e2e %.3e" % s["e2e"]["value"], round(s["roofline"]["frac"], 3), s["config"]["allgather_ms"])
PY
