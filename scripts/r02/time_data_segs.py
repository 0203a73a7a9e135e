"""knn_data filter: one 131072 x 1M x 512 block against the number of reference segments (one process, one data set)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mdsctk_b200, bench
n = int(os.environ.get("N", 1000000))
rows = bench.gen_rows(dict(bench.C5, n_total=n), 0, n)
ctx = mdsctk_b200.KnnContext(0)
ctx.data_set_reference(rows)
for sg in [int(v) for v in os.environ.get("SEGS", "0 1 2 4 8 16").split()]:
    ctx.set_option('data_segments', sg)
    for rep in range(2):
        ctx.data_query(65, metric=0, fetch=False, fit_range=(0, 131072))
    st = ctx.stats()
    print("segments", sg, {k: round(st[k], 2) if isinstance(st[k], float) else st[k] for k in ("ms_sweep", "ms_rescore", "fallback_rows", "lists_per_row")}, flush=True)
