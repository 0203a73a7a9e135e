"""knn_data timings at the C5 shape (reduced rows): Euclidean and correlation on the tensor path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mdsctk_b200, bench
n = int(os.environ.get("N", 200000))
rows = bench.gen_rows(dict(bench.C5, n_total=n), 0, n)
ctx = mdsctk_b200.KnnContext(0)
if os.environ.get('SEGMENTS'):
    ctx.set_option('data_segments', int(os.environ['SEGMENTS']))
ctx.data_set_reference(rows)
if os.environ.get('STREAMING'):
    ctx.set_option('data_streaming', int(os.environ['STREAMING']))
one_block = os.environ.get('ONE_BLOCK') == '1'          # profiling: the first row block only (131072 rows x all reference rows)
for metric in ((0,) if os.environ.get('SEGMENTS') else (0, 1)):
    for rep in range(2):
        ctx.data_query(65, metric=metric, fetch=False, fit_range=(0, 131072) if one_block else None)
    st = ctx.stats()
    tot = st["ms_sweep"] + st["ms_rescore"] + st["ms_fallback"]
    print("metric", metric, {k: round(st[k], 2) if isinstance(st[k], float) else st[k] for k in ("ms_pack", "ms_sweep", "ms_rescore", "ms_fallback", "fallback_rows", "rescored_max", "k_keep", "lists_per_row")},
          "pairs/s %.3e" % ((131072 if one_block else n) * n / tot * 1e3), flush=True)
