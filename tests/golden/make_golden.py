"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference ships no expected outputs and cannot be built in
this image, so these vectors are produced by oracle/ (the restatement of
knn_rms.cpp / knn_data.cpp) on the reference's own example inputs, copied as
data-only fixtures into tests/data/.  They pin the oracle against regressions
and give the GPU tests a target that does not need /root/reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

DATA = os.path.join(ROOT, "tests", "data")
OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print(name, {k: v.shape for k, v in arrs.items()})


def main():
    xyz = ob.read_xtc(os.path.join(DATA, "trp-cage.xtc"))
    mass = ob.read_masses(os.path.join(DATA, "trp-cage.pdb"))
    # decoder pin: a checksum-style digest of the decoded integers + a few frames
    grid = np.round(xyz.astype(np.float64) * 1000).astype(np.int64)
    save("trpcage_decode.npz", mass=mass, frame0=xyz[0], frame999=xyz[999],
         colsum=grid.sum(axis=(1, 2)), total=np.array([grid.sum(), (grid * grid).sum()]))
    for k in (10, 100):
        d1, i1 = ob.knn_rms(xyz, mass, k, mode=1)
        d0, i0 = ob.knn_rms(xyz, mass, k, mode=0)
        save(f"trpcage_rms_k{k}.npz", dist_f64=d1, idx_f64=i1, dist_ref=d0, idx_ref=i0)
    d1, i1 = ob.knn_rms(xyz, mass, 10, mode=1, dofit=False)
    d0, i0 = ob.knn_rms(xyz, mass, 10, mode=0, dofit=False)
    save("trpcage_rms_nofit_k10.npz", dist_f64=d1, idx_f64=i1, dist_ref=d0, idx_ref=i0)
    # out-of-sample shape (-f): fit = every 10th frame, reference = the rest
    fit, ref = xyz[::10], np.delete(xyz, np.arange(0, 1000, 10), axis=0)
    d1, i1 = ob.knn_rms(ref, mass, 10, fit=fit, mode=1)
    save("trpcage_rms_oos_k10.npz", dist_f64=d1, idx_f64=i1)

    for name, dim, ks in (("rings", 2, (10, 20)), ("swissroll", 3, (10, 12))):
        pts = np.fromfile(os.path.join(DATA, f"{name}.pts"), dtype=np.float64)
        pts = pts[: (pts.size // dim) * dim].reshape(-1, dim)
        for k in ks:
            d, i = ob.knn_data(pts, k)
            save(f"{name}_data_k{k}.npz", dist=d, idx=i)
        d, i = ob.knn_data(pts, 10, metric=1) if dim > 2 else (None, None)
        if d is not None:
            save(f"{name}_data_corr_k10.npz", dist=d, idx=i)
    sw = np.fromfile(os.path.join(DATA, "swissroll.pts"), dtype=np.float64).reshape(-1, 3)
    oos = np.fromfile(os.path.join(DATA, "swissroll-outofsample.pts"), dtype=np.float64).reshape(-1, 3)
    d, i = ob.knn_data(sw, 10, fit=oos)
    save("swissroll_data_oos_k10.npz", dist=d, idx=i)


if __name__ == "__main__":
    main()
