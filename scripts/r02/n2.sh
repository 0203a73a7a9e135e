#!/bin/bash
# two-GPU validation: the C++ tools' --gpus path (NCCL inside the tool) and bench.py under torchrun
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2.log
timeout 600 python -m pytest tests/test_tools.py -x -q -m gpu -k "several_gpus" 2>&1 | tail -8 >> gpurun_out/n2.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --secondary c5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 2000 gpurun_out/bench_n2.err >> gpurun_out/n2.log
cat gpurun_out/n2.log
python - <<'PY'
import json
for ln in open("gpurun_out/bench_n2.json"):
    if ln.startswith("{"):
        r = json.loads(ln)
        print({k: r[k] for k in ("value", "ms_per_step", "n_gpus", "lists_hash", "parity")})
        print("e2e", r["e2e"], "roofline", {k: r["roofline"][k] for k in ("achieved", "frac", "sweep_ms_per_step", "step_frac")})
        print("config", {k: r["config"][k] for k in ("allgather_ms", "allgather_bytes_per_gpu", "h2d_ms", "pack_ms", "fallback_rows")})
        for s in r["secondary"]:
            print("secondary", s["config"]["workload"][:30], s["value"], s["lists_hash"], s["parity"], s["e2e"]["value"], s["roofline"]["frac"])
PY
