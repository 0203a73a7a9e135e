"""Iteration helper (GPU box): CSC builder throughput on a C4-sized synthetic kNN graph (n rows, k neighbours,
temporal-locality neighbours like a trajectory's), checked against the oracle on a prefix."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdsctk_b200
from oracle import binding as ob

n = int(os.environ.get("ITER_N", "1000000"))
k = int(os.environ.get("ITER_K", "64"))
rng = np.random.default_rng(5)
# neighbours: mostly close in time (offsets from a wide two-sided geometric law), distinct per row
off = rng.geometric(0.02, size=(n, k)).astype(np.int64) * rng.choice([-1, 1], size=(n, k))
idx = (np.arange(n)[:, None] + np.cumsum(np.abs(off), axis=1) * np.sign(off[:, :1])) % n
idx = idx.astype(np.int32)
dist = np.sort(rng.random((n, k)), axis=1)
ctx = mdsctk_b200.KnnContext(0)
for rep in range(3):
    t = time.time()
    pcol, irow, val = ctx.csc_build_sym(idx, dist)
    wall = time.time() - t
    st = ctx.stats()
edges = n * k
alg_bytes = edges * (12 + 20 + 20) + 12 * int(pcol[-1])
print("csc n=%d k=%d nnz=%d  build %.2f ms (%.2e entries/s, %.0f GB/s algorithmic)  upload %.1f ms  download %.1f ms  wall %.2f s"
      % (n, k, int(pcol[-1]), st["ms_sweep"], edges / st["ms_sweep"] * 1e3, alg_bytes / st["ms_sweep"] / 1e6, st["ms_upload"],
         st["ms_download"], wall), flush=True)
m = min(n, 20000)
t = time.time()
want = ob.make_sysparse(idx[:m] % m, dist[:m])
cpu = time.time() - t
got = ctx.csc_build_sym(idx[:m] % m, dist[:m])
print("oracle check on %d rows:" % m, all(np.array_equal(a, b) for a, b in zip(got, want)), " oracle (qsort, 1 core) %.2e entries/s" % (m * k / cpu))
