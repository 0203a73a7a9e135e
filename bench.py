#!/usr/bin/env python
"""bench.py -- superposed-RMSD frame pairs/sec (all-pairs kNN) on B200, next to the CPU knn_rms.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs, synthetic generators of SURVEY.md section 8d / mdsctk_b200/synth.py):
  N=1   C3: 100 000 frames x 300 atoms, k=32.  One step = one full all-pairs pass
        (1e10 ordered pairs): sweep + FP64 re-score (+ certified fallback rows).
  N>1   C4: 1 000 000 frames x 300 atoms, k=64, row-sharded.  Every rank generates and packs
        ITS shard on ITS GPU, one NCCL all-gather replicates the packed reference set
        (timed once, reported as allgather_ms), then one step = every rank pushes a batch of
        its own fit rows against all 1e6 reference frames (weak scaling: rows per rank per step
        are fixed) -- the same row-block structure as knn_rms.cpp:256-293.  pairs/s does not
        depend on how many row blocks a step holds (per-row cost is constant).

value   whole-job pairs/s with inputs resident in HBM (CUDA events on the library's stream,
        max over ranks).
e2e     the same metric through the public C-ABI call with HOST buffers: H2D of the step's
        frames from pinned memory + pack + sweep + re-score + D2H of the k-lists inside the
        timed region.
roofline  dominant kernel = the sweep; algorithmic flops = 18 * atoms per pair (nine length-A
        dot products) over the sweep's CUDA-event time; peak =
        MEASURED_PEAKS.json's sustained bf16 figure for the 16-bit kernels (the default 1xFP16
        sweep issues one MMA per algorithmic flop, the 3x splits three), half of it for TF32.
cpu_baseline  the oracle's reference-faithful float chain (oracle/, OpenMP, all host cores)
        on the first rows of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "superposed-RMSD frame pairs/sec (all-pairs kNN)"
UNIT = "pairs/s"
ATOMS = 300
FLOP_PER_PAIR = 18 * ATOMS


def workload(n_gpus):
    if n_gpus == 1:
        return dict(name="C3 synthetic trajectory 100k frames x 300 atoms, all-pairs RMSD kNN k=32",
                    n_total=int(os.environ.get("BENCH_FRAMES", 100_000)), k=32, basins=16, seed=20260117,
                    rows_per_rank=None)
    return dict(name="C4 synthetic trajectory 1M frames x 300 atoms, all-pairs RMSD kNN k=64, row-sharded",
                n_total=int(os.environ.get("BENCH_FRAMES", 1_000_000)), k=64, basins=64, seed=20260118,
                rows_per_rank=int(os.environ.get("BENCH_ROWS_PER_RANK", 16384)))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks(rms_kernel):
    """Roofline denominator: MEASURED_PEAKS.json's sustained dense bf16 figure for the bf16 kernel
    (the sweep is timed inside a seconds-long step); half of it for the TF32 / FP32-recovering
    kernels (dense TF32 runs at half the bf16 tensor rate; cuBLAS TF32 8192^3 measured on this pool:
    775 burst / 621 sustained TFLOP/s, profiles/peaks_r01.json)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        m = json.load(open(p))
        bf16, src = m["bf16_tflops_sustained"], "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
    else:
        bf16, src = 1400.0, "fallback ~1.4 PFLOP/s sustained bf16 (B200_PROFILING.md; of fallback)"
    if rms_kernel >= 3:
        return {"tflops": bf16, "source": src + "; kernel issues 16-bit (kind::f16) MMAs"}
    return {"tflops": bf16 / 2.0, "source": src + " / 2 = dense TF32 rate"}


def cpu_baseline_sample(wl, rows=None):
    """Reference-faithful CPU chain (oracle mode 0) on the first `rows` fit rows vs a prefix of the
    reference set sized for ~10-30 s; per-pair cost does not depend on either count."""
    from oracle import binding as ob
    from mdsctk_b200 import synth
    threads = os.cpu_count() or ob.max_threads()   # torchrun exports OMP_NUM_THREADS=1; use every host core
    n_ref = min(wl["n_total"], 20_000)
    rows = rows or 32 * threads
    xyz = synth.traj_frames(wl["n_total"], ATOMS, wl["basins"], wl["seed"], 0, n_ref)
    mass = synth.traj_masses(ATOMS)
    ob.knn_rms(xyz[:2000], mass, 8, fit=xyz[:threads], mode=0, nthreads=threads)  # warm threads
    t = time.perf_counter()
    ob.knn_rms(xyz, mass, wl["k"], fit=xyz[:rows], mode=0, nthreads=threads)
    dt = time.perf_counter() - t
    return {"value": rows * n_ref / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"oracle mode 0 (float do_fit/rmsdev chain restated from GROMACS, OpenMP, -O3): first {rows} fit "
                      f"rows x first {n_ref} reference frames of the same synthetic trajectory, k={wl['k']}, {dt:.1f} s"}


def knn_data_sample(ctx):
    """Second tool on the path (BASELINE.json config 5, reduced to fit the default run): Euclidean
    knn_data on synthetic phi-psi sin/cos rows, k=64, through the same context; not part of `value`."""
    from mdsctk_b200 import synth
    n, dim, k1 = int(os.environ.get("BENCH_DATA_ROWS", 100_000)), 512, 65
    rows = synth.phipsi_rows(n, dim, 64)
    ctx.data_set_reference(rows)
    for _ in range(2):
        ctx.data_query(k1, fetch=False)
    st = ctx.stats()
    tot = st["ms_sweep"] + st["ms_rescore"] + st["ms_fallback"]
    return {"metric": "knn_data Euclidean row pairs/sec (all-pairs kNN)", "value": n * n / tot * 1e3, "unit": "pairs/s",
            "workload": f"synthetic phi-psi sin/cos rows {n} x {dim}, k=64 (C5 shape at reduced row count), 1 GPU",
            "kernel": "data_sweep_tc_kernel (tcgen05 cta_group::2, 1xFP16 operands) + exact FP64 re-score with the rounding term in its certificate, bit-identical output",
            "sweep_ms": st["ms_sweep"], "rescore_ms": st["ms_rescore"], "fallback_rows": st["fallback_rows"],
            "sweep_tflops_algorithmic": n * n * 2 * dim / st["ms_sweep"] * 1e3 / 1e12}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself cannot be
    built here: needs libgromacs/Boost/BDB/ARPACK) on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.gpus)
    from oracle import binding as ob
    from mdsctk_b200 import synth
    threads = os.cpu_count() or ob.max_threads()   # torchrun exports OMP_NUM_THREADS=1; use every host core
    n_ref = min(wl["n_total"], 20_000)
    rows = 8 * threads
    xyz = synth.traj_frames(wl["n_total"], ATOMS, wl["basins"], wl["seed"], 0, n_ref)
    mass = synth.traj_masses(ATOMS)
    for _ in range(args.warmup):
        ob.knn_rms(xyz[:2000], mass, 8, fit=xyz[:threads], mode=0, nthreads=threads)
    t = time.perf_counter()
    for s in range(args.steps):
        ob.knn_rms(xyz, mass, wl["k"], fit=xyz[s * rows:(s + 1) * rows], mode=0, nthreads=threads)
    dt = time.perf_counter() - t
    v = args.steps * rows * n_ref / dt
    sample = f"{rows} fit rows x {n_ref} reference frames per step (oracle mode 0, OpenMP {threads} threads)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "frames": wl["n_total"], "atoms": ATOMS, "k": wl["k"]},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--rms-kernel", type=int, default=int(os.environ.get("BENCH_RMS_KERNEL", "-1")))
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import mdsctk_b200
    from mdsctk_b200 import sharding, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = workload(world)
    n_total, k1 = wl["n_total"], wl["k"] + 1
    mass = synth.traj_masses(ATOMS)
    ctx = mdsctk_b200.KnnContext(local)
    if args.rms_kernel >= 0:
        ctx.set_option("rms_kernel", args.rms_kernel)

    # ---- this rank's frames (pinned host memory) ---------------------------------------------
    begin, count = sharding.shard_range(n_total, world, rank)
    host = torch.empty((count, ATOMS, 3), dtype=torch.float32).pin_memory()
    synth.traj_frames(n_total, ATOMS, wl["basins"], wl["seed"], begin, count, out=host.numpy())

    # ---- resident reference set: pack own shard, all-gather the packed arrays over NCCL ------
    ctx.rms_alloc_reference(n_total, ATOMS, mass)
    ctx.rms_pack_shard(host.numpy(), begin)
    allgather_ms = 0.0
    if world > 1:
        arrays, bpf = ctx.rms_reference_arrays()
        tens = [torch.as_tensor(a, device=f"cuda:{local}") for a in arrays]
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sharding.replicate_frame_major(tens, bpf, n_total, world, rank, dist)
        e1.record()
        torch.cuda.synchronize()
        allgather_ms = e0.elapsed_time(e1)

    if wl["rows_per_rank"] is None:
        rows = count
        fit_range = lambda s: (begin, count)
    else:
        rows = min(wl["rows_per_rank"], count)
        fit_range = lambda s: sharding.step_rows(begin, count, wl["rows_per_rank"], s)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing: W warm-up + K timed steps -------------------------------------
    # the clock sampler (nvidia-smi, 200 ms period) starts before the warm-up so that short timed
    # regions still get samples under load; it is stopped right after the timed steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for s in range(args.warmup):
        ctx.rms_query(k1, fit_range=fit_range(s), fetch=False)
    barrier()
    sweep_ms = rescore_ms = fallback_ms = 0.0
    launches = fallback_rows = 0
    err = 0.0
    ctx.timer_start()
    for s in range(args.steps):
        ctx.rms_query(k1, fit_range=fit_range(args.warmup + s), fetch=False)
        st = ctx.stats()
        sweep_ms += st["ms_sweep"]; rescore_ms += st["ms_rescore"]; fallback_ms += st["ms_fallback"]
        launches += st["launches"]; fallback_rows += st["fallback_rows"]; err = max(err, st["max_filter_err"])
    dev_ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public C-ABI call with host buffers ---------------------------
    e2e_steps = max(1, min(args.steps, 2))
    out_d = torch.empty((rows, k1), dtype=torch.float64).pin_memory()
    out_i = torch.empty((rows, k1), dtype=torch.int32).pin_memory()
    import ctypes as C
    L = mdsctk_b200.load_library()
    dp, ip, fp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float)

    def e2e_step(s):
        if wl["rows_per_rank"] is None:   # C3: upload + pack the whole set, query it against itself
            ctx.rms_set_reference(host.numpy(), mass)
            rc = L.mdsctk_knn_rms_query(ctx._h, None, count, k1, 1, C.cast(out_d.data_ptr(), dp), C.cast(out_i.data_ptr(), ip))
        else:                             # C4: this step's fit rows come from the host
            b, n = fit_range(s)
            src = host[b - begin:b - begin + n]
            rc = L.mdsctk_knn_rms_query(ctx._h, C.cast(src.data_ptr(), fp), n, k1, 1, C.cast(out_d.data_ptr(), dp),
                                        C.cast(out_i.data_ptr(), ip))
        if rc != 0:
            raise RuntimeError(L.mdsctk_knn_last_error(ctx._h).decode())

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        e2e_step(1 + s)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = (count if wl["rows_per_rank"] is None else rows) * ATOMS * 12
    d2h = rows * k1 * 12

    # ---- max over ranks ----------------------------------------------------------------------------
    vec = torch.tensor([dev_ms, e2e_s, sweep_ms, rescore_ms + fallback_ms], dtype=torch.float64, device=f"cuda:{local}")
    tot = torch.tensor([float(launches), float(fallback_rows)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s, sweep_ms, post_ms = vec.tolist()
    launches, fallback_rows = [int(x) for x in tot.tolist()]

    if rank == 0:
        pairs_per_step = world * rows * n_total
        value = pairs_per_step * args.steps / (dev_ms * 1e-3)
        sweep_tflops = world * rows * n_total * FLOP_PER_PAIR * args.steps / (sweep_ms * 1e-3) / 1e12 / world
        st = ctx.stats()
        peaks = measured_peaks(st["rms_kernel"])
        kern = {0: "rms_sweep_simt_kernel (FP32 CUDA-core contraction + QCP + streaming top-k)",
                1: "rms_sweep_tc_kernel<1> (tcgen05 kind::tf32, 3xTF32 split contraction + QCP + streaming top-k)",
                2: "rms_sweep_tc_kernel<2> (tcgen05 kind::tf32, 1xTF32 contraction + QCP + streaming top-k)",
                3: "rms_sweep_tc_kernel<3> (tcgen05 cta_group::2 kind::f16, 3xBF16 split contraction + QCP bounds + streaming top-k)",
                4: "rms_sweep_tc_kernel<4> (tcgen05 cta_group::2 kind::f16, 3xFP16 split contraction + QCP bounds + streaming top-k)",
                5: "rms_sweep_tc_kernel<5> (tcgen05 cta_group::2 kind::f16, 2xFP16 contraction + QCP bounds + streaming top-k)",
                6: "rms_sweep_tc_kernel<6> (tcgen05 cta_group::2 kind::f16, 1xFP16 contraction + QCP bounds + streaming top-k; "
                   "FP64 re-score with the rounded-structure triangle bound)"}[st["rms_kernel"]]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_r01d.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(str(st["rms_kernel"]))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f32", 1: "tf32", 2: "tf32", 3: "bf16", 4: "f16", 5: "f16", 6: "f16"}[st["rms_kernel"]] + " contraction (fp32 accumulate), f64 re-score",
            "data": "synthetic",
            "config": {"workload": wl["name"], "frames": n_total, "atoms": ATOMS, "k": wl["k"],
                       "fit_rows_per_rank_per_step": rows, "parallelism": f"row-sharded x{world}, reference replicated",
                       "l2": "inputs larger than L2 (reference planes %.0f MB + raw %.0f MB vs 126 MB L2)" %
                             (n_total * 3 * 304 * 4 / 1e6, n_total * ATOMS * 12 / 1e6),
                       "kernel": kern, "k_keep": st["k_keep"], "fallback_rows": fallback_rows,
                       "max_filter_err_nm2": err, "cert_eps_nm2": st["cert_eps"], "cert_gres_nm": st["cert_gres"],
                       "rescored_max": st["rescored_max"], "allgather_ms": allgather_ms},
            "e2e": {"value": world * rows * n_total * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": sweep_tflops, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": sweep_tflops / peaks["tflops"], "traffic": traffic,
                         "mma_per_flop": {0: 1, 2: 1, 5: 2, 6: 1}.get(st["rms_kernel"], 3),
                         "kernel": kern.split(" ")[0], "flop_per_pair": FLOP_PER_PAIR, "peak_source": peaks["source"],
                         "sweep_ms_per_step": sweep_ms / args.steps, "post_ms_per_step": post_ms / args.steps},
            "cpu_baseline": cpu_baseline_sample(wl),
        }
        # SURVEY.md section 8d quotes the contraction roofline against the dense TF32 rate (= half the bf16 rate)
        line["roofline"]["frac_of_tf32_rate"] = sweep_tflops / (peaks["tflops"] / 2.0) if st["rms_kernel"] >= 3 else None
        if world == 1 and os.environ.get("BENCH_KNN_DATA", "1") == "1":
            line["secondary"] = knn_data_sample(ctx)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
