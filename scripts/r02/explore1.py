"""Round-2 exploration on the GPU box: timing floors of the existing sweep, C4 whole job at N=1, C5 at full size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import mdsctk_b200
from mdsctk_b200 import synth

what = sys.argv[1]
ctx = mdsctk_b200.KnnContext(0)
if what == "c4":
    n = int(os.environ.get("N", 1000000)); k1 = 65
    t = time.time()
    from concurrent.futures import ThreadPoolExecutor
    xyz = np.empty((n, 300, 3), np.float32)
    parts = 16
    def gen(p):
        b = p * n // parts; e = (p + 1) * n // parts
        synth.traj_frames(n, 300, 64, 20260118, b, e - b, out=xyz[b:e])
    with ThreadPoolExecutor(8) as ex: list(ex.map(gen, range(parts)))
    print("gen s", time.time() - t, flush=True)
    t = time.time(); ctx.rms_set_reference(xyz, synth.traj_masses(300)); print("set_reference s", time.time() - t, ctx.stats()["ms_upload"], ctx.stats()["ms_pack"], flush=True)
    chunk = int(os.environ.get("CHUNK", 131072))
    tot = {"ms_sweep": 0, "ms_rescore": 0, "ms_fallback": 0, "fallback_rows": 0}
    t = time.time()
    for b in range(0, n, chunk):
        c = min(chunk, n - b)
        ctx.rms_query(k1, fit_range=(b, c), fetch=False)
        st = ctx.stats()
        for k in tot: tot[k] += st[k]
        print(b, round(st["ms_sweep"], 1), round(st["ms_rescore"], 1), st["fallback_rows"], st["rescored_max"], flush=True)
    print("C4 N=1 whole job wall s", time.time() - t, tot, "pairs/s %.3e" % (n * n / (tot["ms_sweep"] + tot["ms_rescore"] + tot["ms_fallback"]) * 1e3), flush=True)
elif what == "c5":
    n = int(os.environ.get("N", 1000000)); k1 = 65
    t = time.time(); rows = synth.phipsi_rows(n, 512, 64); print("gen s", time.time() - t, flush=True)
    t = time.time(); ctx.data_set_reference(rows); print("set_reference s", time.time() - t, flush=True)
    chunk = int(os.environ.get("CHUNK", 131072))
    tot = {"ms_sweep": 0, "ms_rescore": 0, "ms_fallback": 0, "fallback_rows": 0, "ms_pack": 0}
    t = time.time()
    for b in range(0, n, chunk):
        c = min(chunk, n - b)
        ctx.data_query(k1, fit_range=(b, c), fetch=False)
        st = ctx.stats()
        for k in tot: tot[k] += st[k]
        print(b, round(st["ms_sweep"], 1), round(st["ms_rescore"], 1), st["fallback_rows"], st["rescored_max"], st["k_keep"], st["lists_per_row"], flush=True)
    print("C5 N=1 whole job wall s", time.time() - t, tot, "pairs/s %.3e" % (n * n / (tot["ms_sweep"] + tot["ms_rescore"] + tot["ms_fallback"]) * 1e3), flush=True)
