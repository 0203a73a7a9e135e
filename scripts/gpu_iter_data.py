"""Iteration helper (GPU box) for the vector path: ITER_N rows x ITER_DIM dims (phi-psi generator), k1."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdsctk_b200
from mdsctk_b200 import synth

n = int(os.environ.get("ITER_N", "200000"))
dim = int(os.environ.get("ITER_DIM", "512"))
k1 = int(os.environ.get("ITER_K1", "65"))
rows_q = int(os.environ.get("ITER_ROWS", "0"))
t = time.time()
rows = synth.phipsi_rows(n, dim, 64)
print("generated", rows.shape, "in %.1f s" % (time.time() - t), flush=True)
ctx = mdsctk_b200.KnnContext(0)
ctx.data_set_reference(rows)
for kern in [int(x) for x in os.environ.get("ITER_KERNELS", "1").split()]:
    ctx.set_option("data_kernel", kern)
    nq = rows_q or n
    if kern == 0:
        nq = min(nq, int(os.environ.get("ITER_EXACT_ROWS", "8192")))
    for rep in range(2):
        ctx.data_query(k1, fit_range=(0, nq), fetch=False)
    st = ctx.stats()
    tot = st["ms_sweep"] + st["ms_rescore"] + st["ms_fallback"]
    print("data_kernel", kern, "rows", nq, {k: (round(st[k], 3) if isinstance(st[k], float) else st[k]) for k in
          ("ms_pack", "ms_sweep", "ms_rescore", "ms_fallback", "fallback_rows", "k_keep", "lists_per_row", "rescored_max")},
          "spread %.2e eps %.2e" % (st["max_filter_spread"], st["cert_eps"]),
          "pairs/s %.3e  sweep-only %.3e  (2*dim flop/pair: %.1f TFLOP/s)" % (nq * n / tot * 1e3, nq * n / st["ms_sweep"] * 1e3,
                                                                             nq * n * 2 * dim / st["ms_sweep"] * 1e3 / 1e12), flush=True)
