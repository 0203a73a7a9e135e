"""ctypes binding for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Never imported by the mdsctk_b200 package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """liboracle.so, and -- where /root/reference is present -- oracle/_ref (the compilable slice of the reference itself)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    subprocess.check_call(["make", "-s", "-C", _HERE, "_ref"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.oracle_xtc_scan.argtypes = [C.c_char_p, ip, C.POINTER(C.c_longlong)]
        L.oracle_xtc_read.argtypes = [C.c_char_p, C.c_int, C.c_longlong, fp]
        L.oracle_top_masses.argtypes = [C.c_char_p, fp, C.c_int]
        L.oracle_reset_x.argtypes = [C.c_int, fp, fp]
        L.oracle_reset_x.restype = None
        L.oracle_do_fit.argtypes = [C.c_int, fp, fp, fp]
        L.oracle_do_fit.restype = None
        L.oracle_rmsdev.argtypes = [C.c_int, fp, fp, fp]
        L.oracle_rmsdev.restype = C.c_float
        L.oracle_rmsd_f64.argtypes = [C.c_int, fp, fp, fp, C.c_int]
        L.oracle_rmsd_f64.restype = C.c_double
        L.oracle_euclidean_distance.argtypes = [C.c_int, dp, dp]
        L.oracle_euclidean_distance.restype = C.c_double
        L.oracle_correlation_distance.argtypes = [C.c_int, dp, dp]
        L.oracle_correlation_distance.restype = C.c_double
        L.oracle_knn_rms.argtypes = [C.c_int, C.c_int, fp, fp, C.c_longlong, fp, C.c_longlong,
                                     C.c_int, C.c_int, C.c_int, dp, ip]
        L.oracle_knn_data.argtypes = [C.c_int, C.c_int, dp, C.c_longlong, dp, C.c_longlong,
                                      C.c_int, C.c_int, dp, ip]
        L.oracle_rms_rows.argtypes = [C.c_int, C.c_int, fp, fp, C.c_longlong, fp, C.c_longlong,
                                      C.c_int, C.c_int, dp]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_make_sysparse.argtypes = [ip, dp, C.c_longlong, C.c_int, C.c_int, ip, ip, dp]
        L.oracle_make_sysparse.restype = C.c_longlong
        L.oracle_make_gesparse.argtypes = [ip, dp, C.c_longlong, C.c_int, C.c_int, C.c_int, ip, ip, dp]
        L.oracle_make_gesparse.restype = C.c_longlong
        lp = C.POINTER(C.c_longlong)
        L.oracle_knn_data_sparse.argtypes = [lp, ip, dp, C.c_longlong, lp, ip, dp, C.c_longlong, C.c_int, dp, ip]
        L.oracle_affinity.argtypes = [C.c_int, ip, ip, dp, C.c_int]
        L.oracle_affinity.restype = C.c_double
        L.oracle_entropic_affinity_sigmas.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp]
        L.oracle_entropic_affinity_sigmas.restype = None
        L.oracle_sym_eigs_largest.argtypes = [C.c_int, ip, ip, dp, C.c_int, dp, dp, dp]
        L.oracle_phipsi.argtypes = [fp, C.c_longlong, C.c_int, dp]
        L.oracle_phipsi.restype = None
        L.oracle_sincos.argtypes = [dp, C.c_longlong, dp]
        L.oracle_sincos.restype = None
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def max_threads():
    return int(lib().oracle_max_threads())


def read_xtc(path):
    """-> float32 [n_frames, n_atoms, 3] in nm (raw, uncentred)."""
    na, nf = C.c_int(0), C.c_longlong(0)
    rc = lib().oracle_xtc_scan(path.encode(), C.byref(na), C.byref(nf))
    if rc != 0:
        raise IOError(f"oracle_xtc_scan({path}) -> {rc}")
    xyz = np.empty((nf.value, na.value, 3), dtype=np.float32)
    rc = lib().oracle_xtc_read(path.encode(), na.value, nf.value, _f(xyz))
    if rc != 0:
        raise IOError(f"oracle_xtc_read({path}) -> {rc}")
    return xyz


def read_masses(path, max_atoms=1 << 20):
    m = np.empty(max_atoms, dtype=np.float32)
    n = lib().oracle_top_masses(path.encode(), _f(m), max_atoms)
    if n < 0:
        raise IOError(f"oracle_top_masses({path}) -> {n}")
    return m[:n].copy()


def reset_x(frame, mass):
    x = np.ascontiguousarray(frame, dtype=np.float32).copy()
    mass = np.ascontiguousarray(mass, dtype=np.float32)
    lib().oracle_reset_x(x.shape[0], _f(x), _f(mass))
    return x


def fit_rmsdev(ref_c, fit_c, mass, dofit=True):
    """Reference float chain on two CENTRED frames; returns (rmsd_nm, fitted copy)."""
    ref_c = np.ascontiguousarray(ref_c, dtype=np.float32)
    fit = np.ascontiguousarray(fit_c, dtype=np.float32).copy()
    mass = np.ascontiguousarray(mass, dtype=np.float32)
    if dofit:
        lib().oracle_do_fit(ref_c.shape[0], _f(mass), _f(ref_c), _f(fit))
    return float(lib().oracle_rmsdev(ref_c.shape[0], _f(mass), _f(ref_c), _f(fit))), fit


def rmsd_f64(a, b, mass, dofit=True):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    mass = np.ascontiguousarray(mass, dtype=np.float32)
    return float(lib().oracle_rmsd_f64(a.shape[0], _f(mass), _f(a), _f(b), int(dofit)))


def euclidean_distance(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return float(lib().oracle_euclidean_distance(a.shape[0], _d(a), _d(b)))


def correlation_distance(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return float(lib().oracle_correlation_distance(a.shape[0], _d(a), _d(b)))


def knn_rms(ref, mass, k, fit=None, mode=1, dofit=True, nthreads=0):
    """ref/fit: float32 [n, atoms, 3] raw frames.  Returns (dist[n_fit,k] A, idx[n_fit,k])."""
    ref = np.ascontiguousarray(ref, dtype=np.float32)
    fit = ref if fit is None else np.ascontiguousarray(fit, dtype=np.float32)
    mass = np.ascontiguousarray(mass, dtype=np.float32)
    k = min(k, ref.shape[0] - 1)
    dist = np.empty((fit.shape[0], k), dtype=np.float64)
    idx = np.empty((fit.shape[0], k), dtype=np.int32)
    rc = lib().oracle_knn_rms(mode, ref.shape[1], _f(mass), _f(ref), ref.shape[0], _f(fit), fit.shape[0],
                              k, int(dofit), nthreads, _d(dist), _i(idx))
    if rc != 0:
        raise RuntimeError(f"oracle_knn_rms -> {rc}")
    return dist, idx


def rms_rows(ref, mass, fit, mode=1, dofit=True, nthreads=0):
    ref = np.ascontiguousarray(ref, dtype=np.float32)
    fit = np.ascontiguousarray(fit, dtype=np.float32)
    mass = np.ascontiguousarray(mass, dtype=np.float32)
    out = np.empty((fit.shape[0], ref.shape[0]), dtype=np.float64)
    rc = lib().oracle_rms_rows(mode, ref.shape[1], _f(mass), _f(ref), ref.shape[0], _f(fit), fit.shape[0],
                               int(dofit), nthreads, _d(out))
    if rc != 0:
        raise RuntimeError(f"oracle_rms_rows -> {rc}")
    return out


def knn_data(ref, k, fit=None, metric=0, nthreads=0):
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    fit = ref if fit is None else np.ascontiguousarray(fit, dtype=np.float64)
    k = min(k, ref.shape[0] - 1)
    dist = np.empty((fit.shape[0], k), dtype=np.float64)
    idx = np.empty((fit.shape[0], k), dtype=np.int32)
    rc = lib().oracle_knn_data(metric, ref.shape[1], _d(ref), ref.shape[0], _d(fit), fit.shape[0],
                               k, nthreads, _d(dist), _i(idx))
    if rc != 0:
        raise RuntimeError(f"oracle_knn_data -> {rc}")
    return dist, idx


def make_sysparse(idx, dist, k=None):
    """(pcol[n+1], irow[nnz], val[nnz]) of the symmetric CSC matrix make_sysparse writes."""
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    dist = np.ascontiguousarray(dist, dtype=np.float64)
    n, maxk = idx.shape
    k = maxk if k is None else k
    pcol = np.zeros(n + 1, dtype=np.int32)
    irow = np.zeros(max(1, n * k), dtype=np.int32)
    val = np.zeros(max(1, n * k), dtype=np.float64)
    nnz = lib().oracle_make_sysparse(idx.ctypes.data_as(C.POINTER(C.c_int)), _d(dist), n, maxk, k,
                                     pcol.ctypes.data_as(C.POINTER(C.c_int)), irow.ctypes.data_as(C.POINTER(C.c_int)), _d(val))
    if nnz < 0:
        raise ValueError("oracle_make_sysparse failed")
    return pcol, irow[:nnz].copy(), val[:nnz].copy()


def make_gesparse(idx, dist, k=None, symmetric=False):
    """(pcol[n+1], irow[nnz], val[nnz]) of the general CSC matrix make_gesparse [-s] writes."""
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    dist = np.ascontiguousarray(dist, dtype=np.float64)
    n, maxk = idx.shape
    k = maxk if k is None else k
    cap = max(1, n * k * (2 if symmetric else 1))
    pcol = np.zeros(n + 1, dtype=np.int32)
    irow = np.zeros(cap, dtype=np.int32)
    val = np.zeros(cap, dtype=np.float64)
    nnz = lib().oracle_make_gesparse(idx.ctypes.data_as(C.POINTER(C.c_int)), _d(dist), n, maxk, k, int(bool(symmetric)),
                                     pcol.ctypes.data_as(C.POINTER(C.c_int)), irow.ctypes.data_as(C.POINTER(C.c_int)), _d(val))
    if nnz < 0:
        raise ValueError("oracle_make_gesparse failed")
    return pcol, irow[:nnz].copy(), val[:nnz].copy()


def phipsi(xyz):
    """double[n][2*(A/3)-2] backbone torsions of N-CA-C frames, as bb_xtc_to_phipsi writes them."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    n, a = xyz.shape[0], xyz.shape[1]
    out = np.empty((n, 2 * (a // 3) - 2), dtype=np.float64)
    lib().oracle_phipsi(_f(xyz), n, a, _d(out))
    return out


def sincos(angles):
    angles = np.ascontiguousarray(angles, dtype=np.float64)
    out = np.empty(angles.size * 2, dtype=np.float64)
    lib().oracle_sincos(_d(angles), angles.size, _d(out))
    return out


def _csr(vectors):
    """[(idx int array, val double array), ...] -> offsets, indices, values."""
    off = np.zeros(len(vectors) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(i) for i, _ in vectors])
    idx = np.concatenate([np.asarray(i, dtype=np.int32) for i, _ in vectors]) if len(vectors) else np.zeros(0, np.int32)
    val = np.concatenate([np.asarray(v, dtype=np.float64) for _, v in vectors]) if len(vectors) else np.zeros(0)
    return off, np.ascontiguousarray(idx, dtype=np.int32), np.ascontiguousarray(val, dtype=np.float64)


def knn_data_sparse(ref, k, fit=None):
    """What knn_data_sparse writes for sparse vectors given as [(indices, values), ...]."""
    fit = ref if fit is None else fit
    ro, ri, rv = _csr(ref)
    fo, fi, fv = _csr(fit)
    k = min(k, len(ref) - 1)
    dist = np.empty((len(fit), k), dtype=np.float64)
    idx = np.empty((len(fit), k), dtype=np.int32)
    lp = C.POINTER(C.c_longlong)
    rc = lib().oracle_knn_data_sparse(ro.ctypes.data_as(lp), ri.ctypes.data_as(C.POINTER(C.c_int)), _d(rv), len(ref),
                                      fo.ctypes.data_as(lp), fi.ctypes.data_as(C.POINTER(C.c_int)), _d(fv), len(fit), k,
                                      _d(dist), idx.ctypes.data_as(C.POINTER(C.c_int)))
    if rc != 0:
        raise ValueError("oracle_knn_data_sparse failed")
    return dist, idx


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def entropic_sigmas(sorted_dist, K):
    """entropic_affinity_sigmas (auto_decomp_sparse -K): sorted_dist [n, k] ascending rows -> sigma[n]."""
    a = np.ascontiguousarray(sorted_dist, dtype=np.float64)
    s = np.empty(a.shape[0])
    lib().oracle_entropic_affinity_sigmas(a.shape[0], a.shape[1], float(K), _d(a), _d(s))
    return s


def affinity(pcol, irow, val, k_a):
    """Normalised affinities of auto_decomp_sparse's affinity stage: (values, average sigma)."""
    pcol = np.ascontiguousarray(pcol, dtype=np.int32)
    irow = np.ascontiguousarray(irow, dtype=np.int32)
    m = np.array(val, dtype=np.float64, copy=True)
    avg = lib().oracle_affinity(len(pcol) - 1, _i(pcol), _i(irow), _d(m), k_a)
    return m, avg


def sym_eigs_largest(pcol, irow, val, nev):
    """(evals[nev] descending, evecs[nev, n], residuals[nev]) of the symmetric upper-CSC matrix (dense Jacobi)."""
    pcol = np.ascontiguousarray(pcol, dtype=np.int32)
    irow = np.ascontiguousarray(irow, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    n = len(pcol) - 1
    ev, vec, res = np.empty(nev), np.empty((nev, n)), np.empty(nev)
    if lib().oracle_sym_eigs_largest(n, _i(pcol), _i(irow), _d(val), nev, _d(ev), _d(vec), _d(res)) != 0:
        raise ValueError("oracle_sym_eigs_largest failed")
    return ev, vec, res
