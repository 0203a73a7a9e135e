"""Out-of-sample query timing (GPU box): fit rows given as host frames vs the same rows addressed as a range of the
reference set; with and without the singular-value start-tile guess (MDSCTK_TC_DEBUG bit 32768 turns it off)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdsctk_b200
from mdsctk_b200 import synth

n = int(os.environ.get("OOS_N", "200000"))
rows, row0, k1 = 16384, 100000, 65
xyz = synth.traj_frames(n, 300, 64, 20260118)
mass = synth.traj_masses(300)
ctx = mdsctk_b200.KnnContext(0)
ctx.rms_set_reference(xyz, mass)
fit = np.ascontiguousarray(xyz[row0:row0 + rows])
for rep in range(2):
    d0, i0 = ctx.rms_query(k1, fit_range=(row0, rows))
st = ctx.stats()
print("in-sample   sweep %.1f rescore %.1f fallback %.1f rows %d" % (st["ms_sweep"], st["ms_rescore"], st["ms_fallback"], st["fallback_rows"]), flush=True)
for dbg in ("0", "32768"):
    os.environ["MDSCTK_TC_DEBUG"] = dbg
    for rep in range(2):
        d1, i1 = ctx.rms_query(k1, fit=fit)
    st = ctx.stats()
    print("out-of-sample dbg %s sweep %.1f rescore %.1f fallback %.1f rows %d identical %s" %
          (dbg, st["ms_sweep"], st["ms_rescore"], st["ms_fallback"], st["fallback_rows"],
           bool(np.array_equal(i0, i1) and np.array_equal(d0, d1))), flush=True)
