"""The oracle pinned against the REFERENCE'S OWN CODE for everything of the path that is not GROMACS.

oracle/_ref/libmdsctk_ref_slice.so is the slice of /root/reference/mdsctk.{h,cpp} that compiles in this image (oracle/ref_slice.sh
cuts the definitions out of the sources where they lie; nothing of it is committed).  Tests marked `needs_slice` call it directly
and are skipped where it is neither built nor buildable; test_oracle_reproduces_the_reference_made_goldens needs only the committed
vectors it wrote (tests/golden/make_ref_slice_golden.py) and runs everywhere, the GPU box included.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402
from oracle import ref_slice as rs  # noqa: E402

DATA = os.path.join(ROOT, "tests", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")
needs_slice = pytest.mark.skipif(not rs.available(), reason="oracle/_ref not built and /root/reference absent")

_spec = importlib.util.spec_from_file_location("make_ref_slice_golden", os.path.join(GOLDEN, "make_ref_slice_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)


def _pts(name, dim):
    pts = np.fromfile(os.path.join(DATA, f"{name}.pts"), dtype=np.float64)
    return pts[: (pts.size // dim) * dim].reshape(-1, dim)


def _phipsi_by(torsion, xyz):
    out = np.empty((xyz.shape[0], 2 * (xyz.shape[1] // 3) - 2))
    for f in range(xyz.shape[0]):
        x = im = 0
        while x < xyz.shape[1] - 3:
            for step in (2, 1):
                out[f, im] = torsion(xyz[f, x], xyz[f, x + 1], xyz[f, x + 2], xyz[f, x + 3], degrees=False)
                im += 1
                x += step
    return out


def test_oracle_reproduces_the_reference_made_goldens():
    """Committed vectors written by the reference's code; the oracle must reproduce every bit.  (No slice needed.)"""
    g = np.load(os.path.join(GOLDEN, "refslice_vectors.npz"))
    for tag, (seed, n, dim, n_fit, k) in {"d64": (11, 300, 64, 40, 12), "d512": (12, 257, 512, 33, 33), "d5": (13, 120, 5, 20, 7)}.items():
        X, F = gen.dense_case(seed, n, dim, n_fit)
        for metric, mn in ((0, "euc"), (1, "cor")):
            d, i = ob.knn_data(X, k, metric=metric)
            assert np.array_equal(d, g[f"{tag}_{mn}_dist"]) and np.array_equal(i, g[f"{tag}_{mn}_idx"]), (tag, mn)
            d, i = ob.knn_data(X, k, fit=F, metric=metric)
            assert np.array_equal(d, g[f"{tag}_{mn}_oos_dist"]) and np.array_equal(i, g[f"{tag}_{mn}_oos_idx"]), (tag, mn, "oos")
    d, i = ob.knn_data_sparse(gen.sparse_case(21, 60, 200), 9)
    assert np.array_equal(d, g["sparse_dist"]) and np.array_equal(i, g["sparse_idx"])
    for tag, (seed, n, k) in {"e30": (31, 500, 30), "e64": (32, 300, 64)}.items():
        A = gen.sorted_rows_case(seed, n, k)
        for K in (3.0, 5.0, 10.0, 17.5):
            assert np.array_equal(ob.entropic_sigmas(A, K), g[f"{tag}_sigma_K{K}"]), (tag, K)
    assert np.array_equal(ob.phipsi(gen.frames_case(41, 50, 30)), g["torsions"])


@needs_slice
def test_committed_goldens_are_what_the_slice_writes_now():
    """The generator is deterministic: re-running it against the reference reproduces the committed file."""
    g = np.load(os.path.join(GOLDEN, "refslice_vectors.npz"))
    X, F = gen.dense_case(12, 257, 512, 33)
    d, i = rs.knn_data(X, 33, fit=F, metric=1)
    assert np.array_equal(d, g["d512_cor_oos_dist"]) and np.array_equal(i, g["d512_cor_oos_idx"])
    assert np.array_equal(rs.entropic_sigmas(gen.sorted_rows_case(31, 500, 30), 5.0), g["e30_sigma_K5.0"])


@needs_slice
def test_the_oracle_made_knn_data_goldens_equal_the_reference_code():
    """tests/golden/{rings,swissroll}_data_*.npz were written by the oracle (make_golden.py) and are what the GPU tests compare
    with, bit for bit; the reference's own distance functions + permutation<double>::sort give exactly the same files."""
    for name, dim, ks in (("rings", 2, (10, 20)), ("swissroll", 3, (10, 12))):
        pts = _pts(name, dim)
        for k in ks:
            g = np.load(os.path.join(GOLDEN, f"{name}_data_k{k}.npz"))
            d, i = rs.knn_data(pts, k)
            assert np.array_equal(d, g["dist"]) and np.array_equal(i, g["idx"]), (name, k)
    sw, oos = _pts("swissroll", 3), _pts("swissroll-outofsample", 3)
    g = np.load(os.path.join(GOLDEN, "swissroll_data_corr_k10.npz"))
    d, i = rs.knn_data(sw, 10, metric=1)
    assert np.array_equal(d, g["dist"]) and np.array_equal(i, g["idx"])
    g = np.load(os.path.join(GOLDEN, "swissroll_data_oos_k10.npz"))
    d, i = rs.knn_data(sw, 10, fit=oos)
    assert np.array_equal(d, g["dist"]) and np.array_equal(i, g["idx"])


@needs_slice
@pytest.mark.parametrize("n,dim,k", [(300, 64, 12), (257, 512, 33), (120, 5, 7), (64, 1, 5), (200, 17, 199)])
def test_dense_metrics_and_selection_bit_identical(n, dim, k):
    rng = np.random.default_rng(n + dim)
    X, F = rng.standard_normal((n, dim)), rng.standard_normal((25, dim))
    for metric in ((0, 1) if dim > 1 else (0,)):
        for fit in (None, F):
            d, i = rs.knn_data(X, k, fit=fit, metric=metric)
            do, io = ob.knn_data(X, k, fit=fit, metric=metric)
            assert np.array_equal(d, do) and np.array_equal(i, io), (metric, fit is None)
    a, b = X[0], X[1]
    assert rs.euclidean_distance(a, b) == ob.euclidean_distance(a, b)
    if dim > 1:                                           # not symmetric in floating point: the order of the arguments matters
        assert rs.correlation_distance(a, b) == ob.correlation_distance(a, b)
        assert rs.correlation_distance(b, a) == ob.correlation_distance(b, a)


@needs_slice
def test_ties_are_the_documented_deviation():
    """std::partial_sort leaves the order of equal distances unspecified (libstdc++: heap order); the oracle and the GPU path
    order them by index.  Same distances, same index SETS per tie group."""
    data = np.array([0.0, 3, 1, 1, 2, 1, 5, 2, 0.5, 1, 7, 1])
    sd, si = rs.partial_sort(data, 7)
    assert np.array_equal(sd, np.sort(data)[:7])
    assert sorted(si[2:7].tolist()) == [2, 3, 5, 9, 11] and si[0] == 0 and si[1] == 8
    X = np.repeat(np.random.default_rng(4).standard_normal((40, 3)), 2, axis=0)      # every row twice
    d, i = rs.knn_data(X, 6)
    do, io = ob.knn_data(X, 6)
    assert np.array_equal(d, do)
    differ = 0
    for r in range(X.shape[0]):
        for j in np.nonzero(i[r] != io[r])[0]:           # a different index only inside a group of equal distances (or on the boundary)
            tied_with_dropped_self = j == 0 and d[r, 0] == 0.0         # the twin row and the row itself: either may be position 0
            assert j == 5 or d[r, j] == d[r, j + 1] or (j > 0 and d[r, j] == d[r, j - 1]) or tied_with_dropped_self, (r, j)
            differ += 1
    print("index differences inside tie groups:", differ)


@needs_slice
def test_sparse_metric_entropic_sigmas_matvec_torsion():
    rng = np.random.default_rng(5)
    vecs = gen.sparse_case(77, 40, 120)
    do, io = ob.knn_data_sparse(vecs, 7)
    for f in range(40):
        row = np.array([rs.euclidean_distance_sparse(vecs[r][0], vecs[r][1], vecs[f][0], vecs[f][1]) for r in range(40)])
        sd, si = rs.partial_sort(row, 8)
        assert np.array_equal(sd[1:], do[f]) and np.array_equal(si[1:], io[f])
    # entropic affinities: bit-identical when the frames' K-th distances are distinct ...
    A = gen.sorted_rows_case(9, 400, 25)
    for K in (3.0, 7.5, 12.0):
        assert np.array_equal(rs.entropic_sigmas(A, K), ob.entropic_sigmas(A, K))
    # ... and within the root finder's tolerance when they tie (a symmetric kNN matrix: mutual neighbours): the reference
    # visits tied frames in std::sort's unspecified order, and each frame starts from its predecessor's solution
    X = rng.standard_normal((400, 6))
    d, _ = ob.knn_data(X, 30)
    assert len(np.unique(d[:, 4])) < 400
    s_ref, s_or = rs.entropic_sigmas(d, 5.0), ob.entropic_sigmas(d, 5.0)
    assert np.max(np.abs(s_ref - s_or) / s_ref) < 1e-8
    # sp_dsymv on a make_sysparse matrix
    d, i = ob.knn_data(rng.standard_normal((300, 4)), 12)
    pc, ir, va = ob.make_sysparse(i, d)
    v = rng.standard_normal(300)
    import ctypes as C
    L = ob.lib()
    w = np.empty(300)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.oracle_sp_dsymv.argtypes = [C.c_int, ip, ip, dp, dp, dp]
    ir32, pc32, va64 = np.ascontiguousarray(ir, np.int32), np.ascontiguousarray(pc, np.int32), np.ascontiguousarray(va, np.float64)
    L.oracle_sp_dsymv(300, ir32.ctypes.data_as(ip), pc32.ctypes.data_as(ip), va64.ctypes.data_as(dp), v.ctypes.data_as(dp), w.ctypes.data_as(dp))
    assert np.array_equal(w, rs.sp_mv(pc, ir, va, v, symmetric=True))
    # torsions with the featuriser's atom pattern (bb_xtc_to_phipsi.cpp:112-121)
    xyz = gen.frames_case(6, 30, 24)
    assert np.array_equal(ob.phipsi(xyz), _phipsi_by(rs.torsion, xyz))
