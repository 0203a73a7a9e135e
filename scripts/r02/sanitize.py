"""Small queries through every default kernel for compute-sanitizer (memcheck / racecheck are too slow for the bench sizes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import mdsctk_b200
from mdsctk_b200 import synth
ctx = mdsctk_b200.KnnContext(0)
for n, atoms in ((700, 300), (1000, 60), (530, 33)):
    xyz = synth.traj_frames(n, atoms, 2, 3)
    d, i = mdsctk_b200.knn_rms(xyz, synth.traj_masses(atoms), 10, ctx=ctx)
    d2, i2 = mdsctk_b200.knn_rms(xyz, synth.traj_masses(atoms), 10, fit_xyz=xyz[::7], ctx=ctx)
    print("rms", n, atoms, ctx.stats()["sweep_version"], d.shape, d2.shape, flush=True)
rows = synth.phipsi_rows(3000, 64, 4)
ctx.set_option("data_kernel", 2)
for corr in (False, True):
    d, i = mdsctk_b200.knn_data(rows, 12, correlation=corr, ctx=ctx)
    print("data", corr, d.shape, ctx.stats()["lists_per_row"], flush=True)
print("done")
