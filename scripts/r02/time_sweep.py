"""C3 / single-basin / C4-block sweep times of the two 1xFP16 sweep generations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import mdsctk_b200
from mdsctk_b200 import synth
import bench
ctx = mdsctk_b200.KnnContext(0)
if os.environ.get('SEGMENTS'):
    ctx.set_option('rms_segments', int(os.environ['SEGMENTS']))
if int(os.environ.get('MDSCTK_TC_DEBUG', '0')):
    ctx.set_option('audit_rows', 0)
ATOMS = int(os.environ.get("ATOMS", 300))
bench.ATOMS = ATOMS
mass = synth.traj_masses(ATOMS)
for name, n, basins, seed, k1, rows in (("C3", 100000, 16, 20260117, 33, 0), ("single-basin", 100000, 1, 20260117, 33, 0),
                                        ("C4 first block", int(os.environ.get("C4N", 1000000)), 64, 20260118, 65, int(os.environ.get("C4ROWS", 132608)))):
    if os.environ.get("ONLY") and os.environ["ONLY"] not in name:
        continue
    xyz = bench.gen_frames(dict(n_total=n, basins=basins, seed=seed), 0, n)
    ctx.rms_set_reference(xyz, mass)
    ref = None
    for ver in [int(v) for v in os.environ.get('VERSIONS', '1 2 2').split()]:
        ctx.set_option("sweep_version", abs(ver))
        ctx.set_option("rms_wide_stages", 0 if ver < 0 else 1)       # -2: version 2 with 32-atom ring stages
        fr = (0, rows) if rows else None
        ctx.rms_query(k1, fit_range=fr, fetch=False)
        d, i = ctx.rms_query(k1, fit_range=fr)
        st = ctx.stats()
        same = None if ref is None else bool(np.array_equal(ref[0], d) and np.array_equal(ref[1], i))
        ref = ref or (d, i)
        print(name, "version", st["sweep_version"], {k: round(st[k], 2) if isinstance(st[k], float) else st[k] for k in
              ("ms_sweep", "ms_rescore", "ms_fallback", "fallback_rows", "rescored_max", "audit_mismatches")}, "identical to v1:", same, flush=True)
