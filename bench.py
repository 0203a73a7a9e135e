#!/usr/bin/env python
"""bench.py -- superposed-RMSD frame pairs/sec (all-pairs kNN) on B200, next to the CPU knn_rms.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload of `value` at EVERY N (BASELINE.json config 4, generators of SURVEY.md section 8d / mdsctk_b200/synth.py):
  C4: 1 000 000 frames x 300 atoms, k=64, the WHOLE job per step: 1e12 ordered pairs, fit rows row-sharded
      across the N ranks, every rank sweeping all N_total/N of its rows against all 1e6 reference frames
      (knn_rms.cpp:256-293) -- strong scaling, the problem is fixed as N grows.
value     1e12 pairs / device time of a step with the packed reference set resident in HBM (CUDA events on the
          library's stream, barrier + synchronize on both sides, max over ranks).
e2e       the same job from HOST frames to HOST k-lists: H2D of the rank's shard from pinned memory, pack,
          NCCL all-gather of the packed arrays, sweep + FP64 re-score of every row, D2H of the lists -- all
          inside the timed region, through the public C-ABI calls.
parity    inside the run: the first rows of rank 0 against the CPU oracle at the full 1M-frame reference set, and a
          partition-independent 64-bit hash of ALL k-lists that must be equal at N = 1, 2, 4, 8.
roofline  dominant kernel = the sweep; algorithmic flops = 18 * atoms per pair (nine length-A dot products, SURVEY.md
          section 8d) over the sweep's CUDA-event time on one GPU; peak = MEASURED_PEAKS.json's sustained bf16 figure
          (the default 1xFP16 sweep issues one 16-bit MMA per algorithmic flop).
cpu_baseline  (N=1) the oracle's reference-faithful float chain (oracle/, OpenMP, all host cores) on the first rows
          of the same workload against the full reference set.
secondary (list) C5 = knn_data 1M x 512, k=64, whole job at the same N, with its own roofline / e2e / parity / hash;
          at N=1 also C3 (100k x 300, k=32: the one-GPU config), its single-basin worst case and the 3xTF32 mode.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "superposed-RMSD frame pairs/sec (all-pairs kNN)"
UNIT = "pairs/s"
ATOMS = 300
FLOP_PER_PAIR = 18 * ATOMS
DIM = 512

C4 = dict(key="c4", name="C4 synthetic trajectory 1M frames x 300 atoms, all-pairs RMSD kNN k=64, whole job, row-sharded",
          n_total=int(os.environ.get("BENCH_FRAMES", 1_000_000)), k=64, basins=64, seed=20260118)
C3 = dict(key="c3", name="C3 synthetic trajectory 100k frames x 300 atoms, all-pairs RMSD kNN k=32, 1 GPU",
          n_total=int(os.environ.get("BENCH_C3_FRAMES", 100_000)), k=32, basins=16, seed=20260117)
C5 = dict(key="c5", name="C5 knn_data synthetic 1M x 512 phi-psi sin/cos rows, Euclidean kNN k=64, whole job, row-sharded",
          n_total=int(os.environ.get("BENCH_DATA_ROWS", 1_000_000)), k=64, basins=64, seed=20260119)

KERNEL_NAMES = {
    0: "rms_sweep_simt_kernel (FP32 CUDA-core contraction + QCP + streaming top-k)",
    1: "rms_sweep_tc_kernel<1> (tcgen05 kind::tf32, 3xTF32 split contraction + QCP + streaming top-k)",
    2: "rms_sweep_tc_kernel<2> (tcgen05 kind::tf32, 1xTF32 contraction + QCP + streaming top-k)",
    3: "rms_sweep_tc_kernel<3> (tcgen05 cta_group::2 kind::f16, 3xBF16 split contraction + QCP bounds + streaming top-k)",
    4: "rms_sweep_tc_kernel<4> (tcgen05 cta_group::2 kind::f16, 3xFP16 split contraction + QCP bounds + streaming top-k)",
    5: "rms_sweep_tc_kernel<5> (tcgen05 cta_group::2 kind::f16, 2xFP16 contraction + QCP bounds + streaming top-k)",
    6: "rms_sweep_tc_kernel<6> (tcgen05 cta_group::2 kind::f16, 1xFP16 contraction + QCP bounds + streaming top-k; "
       "FP64 re-score with the rounded-structure triangle bound)",
}
DTYPES = {0: "f32", 1: "tf32", 2: "tf32", 3: "bf16", 4: "f16", 5: "f16", 6: "f16"}


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
    """Roofline denominator: MEASURED_PEAKS.json's sustained dense bf16 figure for the 16-bit kernels (the sweep is
    timed inside a seconds-long step); half of it for the TF32 / FP32-recovering kernels (dense TF32 runs at half
    the bf16 tensor rate; cuBLAS TF32 8192^3 measured on this pool: 775 burst / 621 sustained, profiles/peaks_r01.json)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        m = json.load(open(p))
        bf16, src = m["bf16_tflops_sustained"], "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
    else:
        bf16, src = 1400.0, "fallback ~1.4 PFLOP/s sustained bf16 (B200_PROFILING.md; of fallback)"
    if rms_kernel >= 3:
        return {"tflops": bf16, "source": src + "; kernel issues 16-bit (kind::f16) MMAs"}
    return {"tflops": bf16 / 2.0, "source": src + " / 2 = dense TF32 rate"}


def measured_traffic(key):
    """DRAM bytes per launch of the dominant kernel from an ncu --set full capture of THIS workload
    (profiles/traffic_r02b.json: the captures of scripts/r02/ncu_r02b.sh), else null."""
    p = os.path.join(ROOT, "profiles", "traffic_r02b.json")
    if os.path.exists(p):
        return json.load(open(p)).get(key)
    return None


def host_threads():
    return os.cpu_count() or 1            # torchrun exports OMP_NUM_THREADS=1; use every host core


def gen_frames(wl, begin, count, out=None, workers=None):
    """Frames [begin, begin+count) of the synthetic trajectory, generated by a thread pool in aligned pieces
    (the generator is indexed by absolute frame number, so the pieces are independent)."""
    from mdsctk_b200 import synth
    if out is None:
        out = np.empty((count, ATOMS, 3), dtype=np.float32)
    piece = 8192 * 4
    cuts = sorted({begin, begin + count} | {x for x in range((begin // piece + 1) * piece, begin + count, piece)})
    jobs = list(zip(cuts[:-1], cuts[1:]))

    def work(j):
        b, e = j
        synth.traj_frames(wl["n_total"], ATOMS, wl["basins"], wl["seed"], b, e - b, out=out[b - begin:e - begin])
    with ThreadPoolExecutor(workers or min(16, host_threads())) as ex:
        list(ex.map(work, jobs))
    return out


def gen_rows(wl, begin, count, out=None, workers=None):
    from mdsctk_b200 import synth
    if out is None:
        out = np.empty((count, DIM), dtype=np.float64)
    piece = 32768
    cuts = sorted({begin, begin + count} | {x for x in range((begin // piece + 1) * piece, begin + count, piece)})

    def work(j):
        b, e = j
        out[b - begin:e - begin] = synth.phipsi_rows(wl["n_total"], DIM, wl["basins"], wl["seed"], b, e - b)
    with ThreadPoolExecutor(workers or min(16, host_threads())) as ex:
        list(ex.map(work, list(zip(cuts[:-1], cuts[1:]))))
    return out


_MULT = None


def lists_hash(dist, idx, row0):
    """Partition-independent 64-bit hash of k-lists: sum over rows of a mix of (global row number, the row's index
    and distance BITS) modulo 2^64 -- equal for every sharding of the same output, sensitive to any changed bit."""
    global _MULT
    k = idx.shape[1]
    if _MULT is None or _MULT.shape[0] < 2 * k:
        _MULT = np.random.default_rng(12345).integers(1, 2 ** 63, size=4 * k, dtype=np.uint64) | np.uint64(1)
    with np.errstate(over="ignore"):
        hi = (idx.astype(np.uint64) * _MULT[None, :k]).sum(axis=1, dtype=np.uint64)
        hd = (np.ascontiguousarray(dist).view(np.uint64) * _MULT[None, k:2 * k]).sum(axis=1, dtype=np.uint64)
        rows = (np.arange(row0, row0 + idx.shape[0], dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        h = (hi ^ (hd * np.uint64(0xC2B2AE3D27D4EB4F))) ^ rows
        h = h * np.uint64(0xD6E8FEB86659FD93)
        h ^= h >> np.uint64(32)
        return int(h.sum(dtype=np.uint64))


class Dist:
    """torch.distributed plumbing (NCCL); no-ops at world size 1."""

    def __init__(self, world, rank, local):
        import torch
        self.torch, self.world, self.rank, self.local = torch, world, rank, local
        self.dev = torch.device("cuda", local)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
            # connection set-up happens on the first collective: keep it out of every timed all-gather
            t = torch.ones(1 << 20, dtype=torch.uint8, device=self.dev)
            g = torch.empty(world << 20, dtype=torch.uint8, device=self.dev)
            dist.all_gather_into_tensor(g, t)
            dist.all_reduce(t)
            torch.cuda.synchronize()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()

    def reduce(self, vals, op):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return t.tolist()

    def sum_u64(self, v):
        t = self.torch.tensor([v & 0xFFFFFFFF, v >> 32], dtype=self.torch.int64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        lo, hi = t.tolist()
        return (lo + (hi << 32)) & 0xFFFFFFFFFFFFFFFF

    def replicate(self, arrays, bpf, n_total):
        """In-place all-gather of frame-major device arrays (mdsctk_b200/sharding.py); returns device ms."""
        if not self.dist:
            return 0.0
        from mdsctk_b200 import sharding
        tens = [self.torch.as_tensor(a, device=self.dev) for a in arrays]
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record()
        sharding.replicate_frame_major(tens, bpf, n_total, self.world, self.rank, self.dist)
        e1.record()
        self.torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
def rms_parity(wl, ctx_dist, ctx_idx, ref_xyz, mass, rows_f64, rows_f32):
    """GPU rows 0.. of rank 0 against the oracle at the FULL reference set: indices identical to the FP64 Kabsch
    (mode 1), distances <= 1e-9 relative; within 1e-4 relative of the reference's float chain (mode 0)."""
    from oracle import binding as ob
    thr = host_threads()
    t = time.perf_counter()
    d1, i1 = ob.knn_rms(ref_xyz, mass, wl["k"], fit=ref_xyz[:rows_f64], mode=1, nthreads=thr)
    gd, gi = ctx_dist[:rows_f64, 1:], ctx_idx[:rows_f64, 1:]
    idx_equal = bool(np.array_equal(gi, i1))
    rel1 = float(np.max(np.abs(gd - d1) / np.maximum(d1, 1e-12)))
    d0, i0 = ob.knn_rms(ref_xyz, mass, wl["k"], fit=ref_xyz[:rows_f32], mode=0, nthreads=thr)
    rel0 = float(np.max(np.abs(ctx_dist[:rows_f32, 1:] - d0) / np.maximum(d0, 1e-12)))
    # how far the reference's float chain itself is from the FP64 answer on the same (row, neighbour) slots: its in-place
    # float rotation of the fit frame accumulates over the whole reference sweep (knn_rms.cpp:272-276), so at 1M
    # frames its own error passes 1e-4 of the smallest distances -- that is a property of the reference, not of the GPU
    same = i0 == i1[:rows_f32]
    chain_own = float(np.max(np.where(same, np.abs(d0 - d1[:rows_f32]) / np.maximum(d1[:rows_f32], 1e-12), 0.0)))
    ok0 = rel0 <= 1e-4 or rel0 <= 1.01 * chain_own + 1e-9
    return {"rows_vs_fp64_kabsch": rows_f64, "indices_identical": idx_equal, "max_rel_dist_err_fp64": rel1,
            "rows_vs_float_chain": rows_f32, "max_rel_dist_err_float_chain": rel0,
            "float_chain_own_rel_err_vs_fp64": chain_own,
            "float_chain_index_mismatches": int((ctx_idx[:rows_f32, 1:] != i0).sum()),
            "ok": bool(idx_equal and rel1 <= 1e-9 and ok0), "reference_frames": int(ref_xyz.shape[0]),
            "tolerance": "indices identical to the FP64 Kabsch oracle and distances within 1e-9 relative of it; within 1e-4 relative "
                         "of the reference's float chain, or within that chain's own deviation from FP64 where it exceeds 1e-4",
            "oracle_s": round(time.perf_counter() - t, 1)}


def run_rms(D, wl, args, rms_kernel=-1, steps=None, warmup=None, want_e2e=True, parity_rows=(128, 32), full=None,
            cpu_baseline=False, sampler=None):
    """One RMSD workload, whole job per step, row-sharded over D.world ranks.  Returns the result dict on rank 0."""
    import torch
    import mdsctk_b200
    from mdsctk_b200 import sharding, synth
    world, rank, local = D.world, D.rank, D.local
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    n_total, k1 = wl["n_total"], wl["k"] + 1
    mass = synth.traj_masses(ATOMS)
    begin, count = sharding.shard_range(n_total, world, rank)
    L = mdsctk_b200.load_library()
    ctx = mdsctk_b200.KnnContext(local)
    if rms_kernel >= 0:
        ctx.set_option("rms_kernel", rms_kernel)

    # rank 0 holds the whole trajectory on the host when it also runs the oracle; every rank holds its shard pinned
    do_parity = rank == 0 and parity_rows[0] > 0
    t0 = time.perf_counter()
    if full is None and (do_parity or cpu_baseline) and rank == 0:
        full = gen_frames(wl, 0, n_total)
    host = torch.empty((count, ATOMS, 3), dtype=torch.float32).pin_memory()
    if full is not None and rank == 0:
        host.numpy()[:] = full[begin:begin + count]
    else:
        gen_frames(wl, begin, count, out=host.numpy())
    gen_s = time.perf_counter() - t0

    out_d = torch.empty((count, k1), dtype=torch.float64).pin_memory()
    out_i = torch.empty((count, k1), dtype=torch.int32).pin_memory()
    dp, ip, fp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float)

    def load_and_replicate():
        """host frames -> packed reference set on every GPU: H2D + pack of the shard, NCCL all-gather"""
        ctx.rms_alloc_reference(n_total, ATOMS, mass)
        ctx.rms_pack_shard(host.numpy(), begin)
        if world == 1:
            return 0.0, 0
        arrays, bpf = ctx.rms_reference_arrays()
        return D.replicate(arrays, bpf, n_total), sum(bpf) * n_total

    allgather_ms, gathered_bytes = load_and_replicate()
    st0 = ctx.stats()
    D.barrier()

    # ---- device-resident timing: W warm-up + K timed steps (a step = the whole job) -----------------------------
    for _ in range(warmup):
        ctx.rms_query(k1, fit_range=(begin, count), fetch=False)
    D.barrier()
    if sampler is not None and rank == 0:
        sampler.start()
    sweep_ms = rescore_ms = fallback_ms = 0.0
    launches = fallback_rows = 0
    err = 0.0
    ctx.timer_start()
    for _ in range(steps):
        ctx.rms_query(k1, fit_range=(begin, count), fetch=False)
        st = ctx.stats()
        sweep_ms += st["ms_sweep"]; rescore_ms += st["ms_rescore"]; fallback_ms += st["ms_fallback"]
        launches += st["launches"]; fallback_rows += st["fallback_rows"]; err = max(err, st["max_filter_err"])
    dev_ms = ctx.timer_stop()
    D.barrier()
    clocks = sampler.stop() if (sampler is not None and rank == 0) else None

    # ---- end to end: host frames -> host k-lists, everything inside ---------------------------------------------
    e2e_s, e2e_steps = 0.0, 0
    if want_e2e:
        e2e_steps = 1 if n_total >= 500_000 else 2

        def e2e_step():
            load_and_replicate()
            rc = L.mdsctk_knn_rms_query_range(ctx._h, begin, count, k1, 1, C.cast(out_d.data_ptr(), dp), C.cast(out_i.data_ptr(), ip))
            if rc != 0:
                raise RuntimeError(L.mdsctk_knn_last_error(ctx._h).decode())
        if n_total < 500_000:
            e2e_step()
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        D.barrier()
        e2e_s = time.perf_counter() - t0
    else:
        ctx.rms_query(k1, fit_range=(begin, count), fetch=False)
        d, i = ctx.fetch(count, k1)
        out_d.numpy()[:] = d; out_i.numpy()[:] = i

    # ---- parity + hash ----------------------------------------------------------------------------------------------
    gd, gi = out_d.numpy(), out_i.numpy()
    h = D.sum_u64(lists_hash(gd[:, 1:], gi[:, 1:], begin))
    self_first = bool((gi[:, 0] == np.arange(begin, begin + count)).mean() > 0.999)
    parity = None
    if do_parity:
        parity = rms_parity(wl, gd, gi, full, mass, min(parity_rows[0], count), min(parity_rows[1], count))
        parity["self_is_rank0"] = self_first
    cpu = None
    if cpu_baseline and rank == 0:
        cpu = cpu_baseline_sample(wl, full, mass)

    dev_ms, e2e_s, sweep_ms, post_ms = D.reduce([dev_ms, e2e_s, sweep_ms, rescore_ms + fallback_ms], "MAX")
    launches, fallback_rows = [int(x) for x in D.reduce([float(launches), float(fallback_rows)], "SUM")]
    st = ctx.stats()
    ctx.close()
    if rank != 0:
        return None
    pairs = float(n_total) * float(n_total)
    kern = st["rms_kernel"]
    kname = KERNEL_NAMES[kern]
    if kern == 6 and st.get("sweep_version") == 2:
        kname = ("rms_sweep_tc2_kernel (tcgen05 cta_group::2 kind::f16, 1xFP16 contraction, fit tile resident in shared memory + TMEM, "
                 "reference-only TMA ring, pass director; QCP bounds + streaming top-k; FP64 re-score with the rounded-structure triangle bound)")
    peaks = measured_peaks(kern)
    sweep_tflops = (count * float(n_total) * FLOP_PER_PAIR * steps) / (sweep_ms * 1e-3) / 1e12      # one GPU's rows / its time
    res = {
        "value": pairs * steps / (dev_ms * 1e-3), "ms_per_step": dev_ms / steps, "steps": steps, "warmup": warmup,
        "dtype": DTYPES[kern] + " contraction (fp32 accumulate), f64 re-score",
        "config": {"workload": wl["name"], "frames": n_total, "atoms": ATOMS, "k": wl["k"],
                   "fit_rows_per_rank_per_step": count, "pairs_per_step": pairs,
                   "parallelism": f"row-sharded x{world}, reference replicated by one NCCL all-gather per array",
                   "l2": "inputs larger than L2 (fp16 reference planes %.0f MB + raw %.0f MB vs 126 MB L2)" %
                         (n_total * 3 * 304 * 2 / 1e6, n_total * ATOMS * 12 / 1e6),
                   "kernel": kname, "k_keep": st["k_keep"], "fallback_rows": fallback_rows,
                   "audit_rows": st["audit_rows"], "audit_mismatches": st["audit_mismatches"],
                   "max_filter_err_nm2": err, "cert_eps_nm2": st["cert_eps"], "cert_gres_nm": st["cert_gres"],
                   "rescored_max": st["rescored_max"], "allgather_ms": allgather_ms, "allgather_bytes_per_gpu": gathered_bytes,
                   "h2d_ms": st0["ms_upload"], "pack_ms": st0["ms_pack"], "host_generation_s": round(gen_s, 1)},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": sweep_tflops, "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": sweep_tflops / peaks["tflops"], "traffic": measured_traffic(wl["key"]),
                     "mma_per_flop": {0: 1, 2: 1, 5: 2, 6: 1}.get(kern, 3), "kernel": kname.split(" ")[0],
                     "flop_per_pair": FLOP_PER_PAIR, "peak_source": peaks["source"],
                     "sweep_ms_per_step": sweep_ms / steps, "post_ms_per_step": post_ms / steps,
                     "step_frac": pairs * FLOP_PER_PAIR * steps / (dev_ms * 1e-3) / 1e12 / world / peaks["tflops"]},
        "lists_hash": "%016x" % h, "parity": parity,
    }
    if want_e2e:
        res["e2e"] = {"value": pairs * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n_total * ATOMS * 12,
                      "d2h_bytes_per_step": n_total * k1 * 12, "steps": e2e_steps,
                      "includes": "H2D of every shard, pack, NCCL all-gather, sweep, re-score, D2H of the k-lists"}
    if cpu is not None:
        res["cpu_baseline"] = cpu
    return res


def cpu_baseline_sample(wl, full, mass, rows=None):
    """Reference-faithful CPU chain (oracle mode 0) on the first `rows` fit rows against the FULL reference set."""
    from oracle import binding as ob
    threads = host_threads()
    rows = rows or 2 * threads
    ob.knn_rms(full[:2000], mass, 8, fit=full[:threads], mode=0, nthreads=threads)  # warm threads
    t = time.perf_counter()
    ob.knn_rms(full, mass, wl["k"], fit=full[:rows], mode=0, nthreads=threads)
    dt = time.perf_counter() - t
    return {"value": rows * full.shape[0] / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"oracle mode 0 (float do_fit/rmsdev chain restated from GROMACS, OpenMP, -O3): first {rows} fit "
                      f"rows x all {full.shape[0]} reference frames of the same synthetic trajectory, k={wl['k']}, {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------------------------
def run_data(D, wl, args, steps=2, warmup=1, parity_rows=256):
    """C5: knn_data whole job, row-sharded; mdsctk_knn_data_alloc_reference / upload_shard / all-gather / query_range."""
    import torch
    import mdsctk_b200
    from mdsctk_b200 import sharding
    world, rank, local = D.world, D.rank, D.local
    n_total, k1 = wl["n_total"], wl["k"] + 1
    begin, count = sharding.shard_range(n_total, world, rank)
    L = mdsctk_b200.load_library()
    ctx = mdsctk_b200.KnnContext(local)
    t0 = time.perf_counter()
    full = gen_rows(wl, 0, n_total) if (rank == 0 and parity_rows > 0) else None
    host = torch.empty((count, DIM), dtype=torch.float64).pin_memory()
    if full is not None:
        host.numpy()[:] = full[begin:begin + count]
    else:
        gen_rows(wl, begin, count, out=host.numpy())
    gen_s = time.perf_counter() - t0
    out_d = torch.empty((count, k1), dtype=torch.float64).pin_memory()
    out_i = torch.empty((count, k1), dtype=torch.int32).pin_memory()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)

    def load_and_replicate():
        ctx.data_alloc_reference(n_total, DIM)
        ctx.data_upload_shard(host.numpy(), begin)
        if world == 1:
            return 0.0
        arrays, bpf = ctx.data_reference_arrays()
        return D.replicate(arrays, bpf, n_total)

    allgather_ms = load_and_replicate()
    D.barrier()
    for _ in range(warmup):
        ctx.data_query(k1, fit_range=(begin, count), fetch=False)
    D.barrier()
    sweep_ms = post_ms = 0.0
    launches = fallback_rows = 0
    ctx.timer_start()
    for _ in range(steps):
        ctx.data_query(k1, fit_range=(begin, count), fetch=False)
        st = ctx.stats()
        sweep_ms += st["ms_sweep"]; post_ms += st["ms_rescore"] + st["ms_fallback"]
        launches += st["launches"]; fallback_rows += st["fallback_rows"]
    dev_ms = ctx.timer_stop()
    D.barrier()

    def e2e_step():
        load_and_replicate()
        rc = L.mdsctk_knn_data_query_range(ctx._h, begin, count, k1, 0, C.cast(out_d.data_ptr(), dp), C.cast(out_i.data_ptr(), ip))
        if rc != 0:
            raise RuntimeError(L.mdsctk_knn_last_error(ctx._h).decode())
    D.barrier()
    t0 = time.perf_counter()
    e2e_step()
    D.barrier()
    e2e_s = time.perf_counter() - t0

    gd, gi = out_d.numpy(), out_i.numpy()
    h = D.sum_u64(lists_hash(gd[:, 1:], gi[:, 1:], begin))
    parity = cpu = None
    if full is not None:
        from oracle import binding as ob
        thr = host_threads()
        r = min(parity_rows, count)
        t = time.perf_counter()
        d, i = ob.knn_data(full, wl["k"], fit=full[:r], nthreads=thr)
        dt = time.perf_counter() - t
        parity = {"rows_vs_oracle": r, "reference_rows": n_total, "bit_identical_distances": bool(np.array_equal(gd[:r, 1:], d)),
                  "indices_identical": bool(np.array_equal(gi[:r, 1:], i)), "oracle_s": round(dt, 1)}
        parity["ok"] = parity["bit_identical_distances"] and parity["indices_identical"]
        cpu = {"value": r * n_total / dt, "unit": UNIT, "cores": thr, "kind": "port",
               "sample": f"oracle knn_data (sequential double sum of squares, mdsctk.cpp:330-335, OpenMP): first {r} fit rows x all "
                         f"{n_total} reference rows, {dt:.1f} s"}
        # the reference's OWN distance function + permutation sort (oracle/_ref: the slice of mdsctk.cpp / mdsctk.h that compiles,
        # built from /root/reference in the container and shipped as a .so) on a smaller sample of the same rows; a checker and a
        # baseline like the oracle above, and never fatal: the oracle check stands on its own
        try:
            from oracle import ref_slice as rs
            if os.path.exists(rs.SO):                            # prebuilt by build(); never built (nor /root/reference looked at) here
                rr = min(64, r)
                t = time.perf_counter()
                d2, i2 = rs.knn_data(full, wl["k"], fit=full[:rr], nthreads=thr)
                dt2 = time.perf_counter() - t
                parity["rows_vs_reference_code"] = rr
                parity["bit_identical_to_reference_code"] = bool(np.array_equal(gd[:rr, 1:], d2) and np.array_equal(gi[:rr, 1:], i2))
                parity["ok"] = parity["ok"] and parity["bit_identical_to_reference_code"]
                cpu = {"value": rr * n_total / dt2, "unit": UNIT, "cores": thr, "kind": "reference",
                       "sample": f"the reference's euclidean_distance + permutation<double>::sort (mdsctk.cpp:330-335, mdsctk.h:177-199, compiled "
                                 f"unmodified: oracle/ref_slice.sh) under the row loop of knn_data.cpp:195-250, OpenMP: first {rr} fit rows x all "
                                 f"{n_total} reference rows, {dt2:.1f} s (oracle port on {r} rows: {r * n_total / dt:.3g} pairs/s)"}
        except Exception as e:                                   # noqa: BLE001
            parity["reference_code_error"] = str(e)[:200]
    dev_ms, e2e_s, sweep_ms, post_ms = D.reduce([dev_ms, e2e_s, sweep_ms, post_ms], "MAX")
    launches, fallback_rows = [int(x) for x in D.reduce([float(launches), float(fallback_rows)], "SUM")]
    st = ctx.stats()
    ctx.close()
    if rank != 0:
        return None
    pairs = float(n_total) * float(n_total)
    peaks = measured_peaks(6)
    sweep_tflops = count * float(n_total) * 2 * DIM * steps / (sweep_ms * 1e-3) / 1e12
    return {
        "metric": "knn_data Euclidean row pairs/sec (all-pairs kNN)", "value": pairs * steps / (dev_ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": dev_ms / steps, "scaling": "strong", "dtype": "f16 contraction (fp32 accumulate), f64 re-score",
        "config": {"workload": wl["name"], "rows": n_total, "dim": DIM, "k": wl["k"], "fit_rows_per_rank_per_step": count,
                   "kernel": "data_sweep_tc_kernel (tcgen05 cta_group::2, 1xFP16 operands) + exact FP64 re-score in the reference's operation order",
                   "k_keep": st["k_keep"], "fallback_rows": fallback_rows, "allgather_ms": allgather_ms, "host_generation_s": round(gen_s, 1)},
        "e2e": {"value": pairs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n_total * DIM * 8, "d2h_bytes_per_step": n_total * k1 * 12,
                "includes": "H2D of every shard, NCCL all-gather, pack, sweep, re-score, D2H of the k-lists"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": sweep_tflops, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": sweep_tflops / peaks["tflops"],
                     "traffic": measured_traffic(wl["key"]), "kernel": "data_sweep_tc_kernel", "flop_per_pair": 2 * DIM,
                     "peak_source": peaks["source"], "sweep_ms_per_step": sweep_ms / steps, "post_ms_per_step": post_ms / steps,
                     "step_frac": pairs * 2 * DIM * steps / (dev_ms * 1e-3) / 1e12 / world / peaks["tflops"]},
        "cpu_baseline": cpu, "lists_hash": "%016x" % h, "parity": parity,
    }


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself cannot be built here: it
    needs libgromacs / Boost / Berkeley DB / ARPACK) on the host cores.  Same config as the B200 arm (C4), each step
    a bounded sample: one fit row per host thread against the FULL 1M-frame reference set."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = C4
    from oracle import binding as ob
    from mdsctk_b200 import synth
    threads = host_threads()
    n_ref = wl["n_total"]
    rows = threads
    xyz = gen_frames(wl, 0, n_ref)
    mass = synth.traj_masses(ATOMS)
    for w in range(min(args.warmup, 2)):
        ob.knn_rms(xyz, mass, wl["k"], fit=xyz[w * threads:(w + 1) * threads][:max(1, threads // 4)], mode=0, nthreads=threads)
    t = time.perf_counter()
    for s in range(args.steps):
        ob.knn_rms(xyz, mass, wl["k"], fit=xyz[s * rows:(s + 1) * rows], mode=0, nthreads=threads)
    dt = time.perf_counter() - t
    v = args.steps * rows * n_ref / dt
    sample = f"{rows} fit rows x all {n_ref} reference frames per step (oracle mode 0, OpenMP {threads} threads)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "frames": wl["n_total"], "atoms": ATOMS, "k": wl["k"], "n_ref": n_ref,
                   "fit_rows_per_step": rows},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--rms-kernel", type=int, default=int(os.environ.get("BENCH_RMS_KERNEL", "-1")))
    ap.add_argument("--secondary", default=os.environ.get("BENCH_SECONDARY", "c5,c3,c3_single_basin,c3_3xtf32"),
                    help="comma list of extra workloads reported under 'secondary' (c3* only at N=1); 'none' to skip")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    D = Dist(world, rank, local)
    sec = [] if args.secondary == "none" else [s.strip() for s in args.secondary.split(",") if s.strip()]

    sampler = ClockSampler(local)
    main_res = run_rms(D, C4, args, rms_kernel=args.rms_kernel, cpu_baseline=(world == 1), sampler=sampler,
                       parity_rows=(int(os.environ.get("BENCH_PARITY_ROWS", 128)), int(os.environ.get("BENCH_PARITY_ROWS_F32", 32))))
    secondary = []
    if "c5" in sec:
        r = run_data(D, C5, args)
        if r:
            secondary.append(r)
    if world == 1:
        s_steps = min(args.steps, 5)
        if "c3" in sec:
            r = run_rms(D, C3, args, steps=s_steps, warmup=3, parity_rows=(256, 64))
            r.update({"metric": METRIC, "unit": UNIT, "n_gpus": 1, "scaling": "n/a (one-GPU config)"})
            secondary.append(r)
        if "c3_single_basin" in sec:
            wl = dict(C3, key="c3_single_basin", basins=1,
                      name="C3 shape, ONE conformational basin (worst case: no pair is far, every accumulator batch is read)")
            r = run_rms(D, wl, args, steps=2, warmup=2, want_e2e=False, parity_rows=(64, 16))
            r.update({"metric": METRIC, "unit": UNIT, "n_gpus": 1})
            secondary.append(r)
        if "c3_3xtf32" in sec:
            wl = dict(C3, key="c3_3xtf32", name="C3 with the 3xTF32 precision-recovering contraction north_star names (rms_kernel 1)")
            r = run_rms(D, wl, args, rms_kernel=1, steps=2, warmup=2, want_e2e=False, parity_rows=(64, 16))
            r.update({"metric": METRIC, "unit": UNIT, "n_gpus": 1})
            secondary.append(r)
    if rank == 0:
        line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": main_res["steps"],
                "warmup": main_res["warmup"], "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "data": "synthetic"}
        for k in ("dtype", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "lists_hash", "parity"):
            if k in main_res:
                line[k] = main_res[k]
        line["secondary"] = secondary
        print(json.dumps(line))
    D.close()


if __name__ == "__main__":
    main()
