"""Iteration helper (GPU box): one process, one synthetic trajectory, several MDSCTK_TC_DEBUG settings.
   ITER_N frames, ITER_DBG="0 1 5 3", ITER_BASINS, ITER_KERNEL; prints one line per setting."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdsctk_b200
from mdsctk_b200 import synth

n = int(os.environ.get("ITER_N", "100000"))
k1 = int(os.environ.get("ITER_K1", "33"))
xyz = synth.traj_frames(n, 300, int(os.environ.get("ITER_BASINS", "16")), int(os.environ.get("ITER_SEED", "20260117")))
rows = int(os.environ.get("ITER_ROWS", "0"))
if os.environ.get("ITER_SLACK"):
    pass
mass = synth.traj_masses(300)
ctx = mdsctk_b200.KnnContext(0)
if os.environ.get("ITER_CERT_PPM"):
    ctx.set_option("cert_scale_ppm", int(os.environ["ITER_CERT_PPM"]))
if os.environ.get("ITER_SLACK"):
    ctx.set_option("slack", int(os.environ["ITER_SLACK"]))
ctx.rms_set_reference(xyz, mass)
for spec in os.environ.get("ITER_DBG", "0 1").split():
    kern, _, dbg = spec.rpartition(":")
    kern = int(kern) if kern else int(os.environ.get("ITER_KERNEL", "3"))
    ctx.set_option("rms_kernel", kern)
    os.environ["MDSCTK_TC_DEBUG"] = dbg
    try:
        for rep in range(int(os.environ.get("ITER_REPS", "2"))):
            ctx.rms_query(k1, fetch=False, fit_range=(int(os.environ.get("ITER_ROW0", "0")), rows) if rows else None)
        st = ctx.stats()
        import time as _t
        ctx.timer_start(); ctx.rms_query(k1, fetch=False, fit_range=(int(os.environ.get("ITER_ROW0", "0")), rows) if rows else None); tot_ms = ctx.timer_stop()
        print("step_ms %.2f" % tot_ms, end=" ")
        print("kernel", kern, "dbg", dbg, {k: (round(st[k], 3) if isinstance(st[k], float) else st[k]) for k in
                           ("ms_sweep", "ms_rescore", "ms_fallback", "fallback_rows", "lists_per_row", "k_keep", "rescored_max")},
              "spread %.2e err %.2e eps %.2e" % (st["max_filter_spread"], st["max_filter_err"], st["cert_eps"]),
              "pairs/s %.3e" % ((rows or n) * n / (st["ms_sweep"] + st["ms_rescore"] + st["ms_fallback"]) * 1e3), flush=True)
    except Exception as e:  # noqa: BLE001
        print("dbg", dbg, "FAILED", e, flush=True)
        break
