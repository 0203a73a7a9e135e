"""ctypes doors into oracle/_ref/libmdsctk_ref_slice.so -- the part of the REFERENCE ITSELF that compiles here (the
arithmetic of knn_data, the selection, the sparse metric, the entropic affinities, the sparse matrix-vector products and the
torsion of mdsctk.cpp / mdsctk.h, cut out of /root/reference by oracle/ref_slice.sh at build time and compiled unmodified).

Test infrastructure: it pins the oracle's restatement (tests/test_ref_slice.py) and writes golden vectors
(tests/golden/make_ref_slice_golden.py).  Never imported by the product, bench.py or smoke()."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libmdsctk_ref_slice.so")
_lib = None


def build():
    """Runs oracle/ref_slice.sh (a no-op where /root/reference is absent).  Returns True when the library exists."""
    subprocess.run(["bash", os.path.join(HERE, "ref_slice.sh")], check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(SO)


def available():
    return os.path.exists(SO) or build()


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libmdsctk_ref_slice.so is not built and /root/reference is absent")
        L = C.CDLL(SO)
        dp, ip, fp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float)
        L.ref_euclidean_distance.restype = L.ref_correlation_distance.restype = L.ref_euclidean_distance_sparse.restype = C.c_double
        L.ref_euclidean_distance.argtypes = L.ref_correlation_distance.argtypes = [C.c_int, dp, dp]
        L.ref_euclidean_distance_sparse.argtypes = [C.c_int, ip, dp, C.c_int, ip, dp]
        L.ref_partial_sort.argtypes = [C.c_int, C.c_int, dp, dp, ip]
        L.ref_knn_data_rows.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, ip, C.c_int]
        L.ref_entropic_affinity_sigmas.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp]
        L.ref_sp_dsymv.argtypes = L.ref_sp_dgemv.argtypes = [C.c_int, ip, ip, dp, dp, dp]
        L.ref_torsion.restype = C.c_float
        L.ref_torsion.argtypes = [fp, fp, fp, fp, C.c_int]
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def euclidean_distance(a, b):
    a, b = _d(a), _d(b)
    return lib().ref_euclidean_distance(a.size, _p(a, C.c_double), _p(b, C.c_double))


def correlation_distance(a, b):
    a, b = _d(a), _d(b)
    return lib().ref_correlation_distance(a.size, _p(a, C.c_double), _p(b, C.c_double))


def euclidean_distance_sparse(ri, rd, fi, fd):
    ri, rd, fi, fd = _i(ri), _d(rd), _i(fi), _d(fd)
    return lib().ref_euclidean_distance_sparse(ri.size, _p(ri, C.c_int), _p(rd, C.c_double), fi.size, _p(fi, C.c_int), _p(fd, C.c_double))


def partial_sort(data, k):
    data = _d(data)
    m = k if k else data.size
    out, idx = np.empty(m), np.empty(m, np.int32)
    lib().ref_partial_sort(data.size, k, _p(data, C.c_double), _p(out, C.c_double), _p(idx, C.c_int))
    return out, idx


def knn_data(ref, k, fit=None, metric=0, nthreads=0):
    ref = _d(ref)
    fit = ref if fit is None else _d(fit)
    dist, idx = np.empty((fit.shape[0], k)), np.empty((fit.shape[0], k), np.int32)
    lib().ref_knn_data_rows(fit.shape[0], ref.shape[0], ref.shape[1], k, metric, _p(fit, C.c_double), _p(ref, C.c_double),
                            _p(dist, C.c_double), _p(idx, C.c_int), int(nthreads))
    return dist, idx


def entropic_sigmas(sorted_dist, K):
    A = _d(sorted_dist)
    s = np.empty(A.shape[0])
    lib().ref_entropic_affinity_sigmas(A.shape[0], A.shape[1], float(K), _p(A, C.c_double), _p(s, C.c_double))
    return s


def sp_mv(pcol, irow, val, v, symmetric=True):
    pcol, irow, val, v = _i(pcol), _i(irow), _d(val), _d(v)
    w = np.empty(v.size)
    fn = lib().ref_sp_dsymv if symmetric else lib().ref_sp_dgemv
    fn(v.size, _p(irow, C.c_int), _p(pcol, C.c_int), _p(val, C.c_double), _p(v, C.c_double), _p(w, C.c_double))
    return w


def torsion(p1, p2, p3, p4, degrees=True):
    ps = [np.ascontiguousarray(p, dtype=np.float32) for p in (p1, p2, p3, p4)]
    return float(lib().ref_torsion(*[_p(p, C.c_float) for p in ps], int(degrees)))
