"""Golden vectors produced by the REFERENCE'S OWN CODE (oracle/_ref: the slice of /root/reference/mdsctk.{h,cpp} that
compiles in this image, see oracle/ref_slice.sh) on seeded inputs.  They travel to the GPU box, where /root/reference does not
exist, and pin the oracle (tests/test_ref_slice.py::test_oracle_reproduces_the_reference_made_goldens) and through it every
GPU parity test of the vector path.  Inputs are regenerated from the seeds below, only outputs (and small inputs) are stored.

    python tests/golden/make_ref_slice_golden.py        # needs /root/reference
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_slice as rs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def dense_case(seed, n, dim, n_fit):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, dim)), rng.standard_normal((n_fit, dim))


def sparse_case(seed, n, dimn):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        nnz = int(rng.integers(1, 40))
        out.append((np.sort(rng.choice(dimn, nnz, replace=False)).astype(np.int32), rng.standard_normal(nnz)))
    return out


def sorted_rows_case(seed, n, k):
    rng = np.random.default_rng(seed)
    return np.sort(rng.random((n, k)) + 0.05, axis=1)          # no ties in any column (the reference's frame order is then unique)


def frames_case(seed, n, atoms):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, atoms, 3)) * 0.3).astype(np.float32)


def main():
    assert rs.available(), "oracle/_ref is not built (needs /root/reference)"
    out = {}
    for tag, (seed, n, dim, n_fit, k) in {"d64": (11, 300, 64, 40, 12), "d512": (12, 257, 512, 33, 33), "d5": (13, 120, 5, 20, 7)}.items():
        X, F = dense_case(seed, n, dim, n_fit)
        for metric, mn in ((0, "euc"), (1, "cor")):
            d, i = rs.knn_data(X, k, metric=metric)
            out[f"{tag}_{mn}_dist"], out[f"{tag}_{mn}_idx"] = d, i
            d, i = rs.knn_data(X, k, fit=F, metric=metric)
            out[f"{tag}_{mn}_oos_dist"], out[f"{tag}_{mn}_oos_idx"] = d, i
    vecs = sparse_case(21, 60, 200)
    dist = np.empty((60, 9)); idx = np.empty((60, 9), np.int32)
    for f in range(60):
        row = np.array([rs.euclidean_distance_sparse(vecs[r][0], vecs[r][1], vecs[f][0], vecs[f][1]) for r in range(60)])
        sd, si = rs.partial_sort(row, 10)
        dist[f], idx[f] = sd[1:], si[1:]
    out["sparse_dist"], out["sparse_idx"] = dist, idx
    for tag, (seed, n, k) in {"e30": (31, 500, 30), "e64": (32, 300, 64)}.items():
        A = sorted_rows_case(seed, n, k)
        for K in (3.0, 5.0, 10.0, 17.5):
            out[f"{tag}_sigma_K{K}"] = rs.entropic_sigmas(A, K)
    xyz = frames_case(41, 50, 30)
    tors = np.empty((50, 18))
    for f in range(50):
        x = im = 0
        while x < 30 - 3:                                       # bb_xtc_to_phipsi.cpp:112-121: steps of 2 and 1 along N-CA-C
            for step in (2, 1):
                tors[f, im] = rs.torsion(xyz[f, x], xyz[f, x + 1], xyz[f, x + 2], xyz[f, x + 3], degrees=False)
                im += 1
                x += step
    out["torsions"] = tors
    np.savez_compressed(os.path.join(OUT, "refslice_vectors.npz"), **out)
    print("refslice_vectors.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
