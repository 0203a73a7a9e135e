"""ctypes binding of include/mdsctk_knn.h.

``knn_rms`` / ``knn_data`` mirror the reference tools' row loop + writer
(knn_rms.cpp:231-293, knn_data.cpp:195-250): they ask the library for k+1 neighbours and
drop sorted position 0, returning exactly the arrays the tools write to distances.dat /
indices.dat.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RMS_SIMT_FP32, RMS_TC_3XTF32, RMS_TC_1XTF32, RMS_TC_3XBF16, RMS_TC_3XFP16, RMS_TC_2XFP16, RMS_TC_1XFP16 = 0, 1, 2, 3, 4, 5, 6
EUCLIDEAN, CORRELATION = 0, 1

SYMBOLS = [
    "mdsctk_knn_abi_version", "mdsctk_knn_create", "mdsctk_knn_destroy", "mdsctk_knn_last_error",
    "mdsctk_knn_set_option", "mdsctk_knn_get_stats",
    "mdsctk_knn_rms_set_reference", "mdsctk_knn_rms_query", "mdsctk_knn_rms_alloc_reference",
    "mdsctk_knn_rms_pack_shard", "mdsctk_knn_rms_reference_arrays", "mdsctk_knn_rms_query_range",
    "mdsctk_knn_data_set_reference", "mdsctk_knn_data_query", "mdsctk_knn_data_alloc_reference",
    "mdsctk_knn_data_upload_shard", "mdsctk_knn_data_reference_arrays", "mdsctk_knn_data_query_range",
    "mdsctk_knn_fetch", "mdsctk_knn_rms_rows", "mdsctk_knn_timer_start", "mdsctk_knn_timer_stop",
    "mdsctk_knn_debug_fetch_tile", "mdsctk_knn_csc_build_sym", "mdsctk_knn_csc_build_general", "mdsctk_knn_csc_fetch",
    "mdsctk_knn_phipsi", "mdsctk_knn_sincos", "mdsctk_knn_data_rows", "mdsctk_knn_spectral_decomp",
    "mdsctk_knn_debug_fetch_array", "mdsctk_knn_debug_rms_layout", "mdsctk_knn_spectral_decomp_ex",
]


class KnnError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [("ms_upload", C.c_double), ("ms_pack", C.c_double), ("ms_sweep", C.c_double),
                ("ms_rescore", C.c_double), ("ms_fallback", C.c_double), ("ms_download", C.c_double),
                ("pairs", C.c_longlong), ("launches", C.c_longlong), ("fallback_rows", C.c_longlong),
                ("sweep_appends", C.c_longlong), ("max_filter_err", C.c_double), ("max_filter_spread", C.c_double),
                ("cert_eps", C.c_double), ("rms_kernel", C.c_int), ("k_keep", C.c_int), ("lists_per_row", C.c_int),
                ("rescored_max", C.c_int), ("cert_gres", C.c_double), ("audit_rows", C.c_longlong),
                ("audit_mismatches", C.c_longlong), ("sweep_version", C.c_int)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def library_path():
    # MDSCTK_KNN_LIBRARY: an alternative build of the same sources (e.g. the clock-counter build of
    # scripts/build_prof.sh); never a different implementation
    return os.environ.get("MDSCTK_KNN_LIBRARY") or os.path.join(_PKG, "libmdsctk_knn.so")


def load_library():
    """Loads libmdsctk_knn.so.  Raises if it has not been built -- there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise KnnError(f"{path} is missing: run `python -m mdsctk_b200.build` (nvcc, sm_100a). "
                       "mdsctk_b200 has no CPU fallback.")
    L = C.CDLL(path)
    vp, fp, dp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
    ll = C.c_longlong
    L.mdsctk_knn_abi_version.restype = C.c_int
    L.mdsctk_knn_create.argtypes = [C.POINTER(vp), C.c_int]
    L.mdsctk_knn_destroy.argtypes = [vp]
    L.mdsctk_knn_destroy.restype = None
    L.mdsctk_knn_last_error.argtypes = [vp]
    L.mdsctk_knn_last_error.restype = C.c_char_p
    L.mdsctk_knn_set_option.argtypes = [vp, C.c_char_p, ll]
    L.mdsctk_knn_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.mdsctk_knn_rms_set_reference.argtypes = [vp, fp, ll, C.c_int, fp]
    L.mdsctk_knn_rms_query.argtypes = [vp, fp, ll, C.c_int, C.c_int, dp, ip]
    L.mdsctk_knn_rms_alloc_reference.argtypes = [vp, ll, C.c_int, fp]
    L.mdsctk_knn_rms_pack_shard.argtypes = [vp, fp, ll, ll]
    L.mdsctk_knn_rms_reference_arrays.argtypes = [vp, C.c_int, ip, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mdsctk_knn_rms_query_range.argtypes = [vp, ll, ll, C.c_int, C.c_int, dp, ip]
    L.mdsctk_knn_data_set_reference.argtypes = [vp, dp, ll, C.c_int]
    L.mdsctk_knn_data_query.argtypes = [vp, dp, ll, C.c_int, C.c_int, dp, ip]
    L.mdsctk_knn_data_alloc_reference.argtypes = [vp, ll, C.c_int]
    L.mdsctk_knn_data_upload_shard.argtypes = [vp, dp, ll, ll]
    L.mdsctk_knn_data_reference_arrays.argtypes = [vp, C.c_int, ip, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.mdsctk_knn_data_query_range.argtypes = [vp, ll, ll, C.c_int, C.c_int, dp, ip]
    L.mdsctk_knn_fetch.argtypes = [vp, dp, ip]
    L.mdsctk_knn_rms_rows.argtypes = [vp, ll, ll, C.c_int, dp]
    L.mdsctk_knn_timer_start.argtypes = [vp]
    L.mdsctk_knn_debug_fetch_tile.argtypes = [vp, fp]
    L.mdsctk_knn_timer_stop.argtypes = [vp, dp]
    L.mdsctk_knn_debug_fetch_array.argtypes = [vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.mdsctk_knn_debug_rms_layout.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.mdsctk_knn_csc_build_sym.argtypes = [vp, ip, dp, ll, C.c_int, C.c_int, ip, C.POINTER(ll)]
    L.mdsctk_knn_csc_build_general.argtypes = [vp, ip, dp, ll, C.c_int, C.c_int, C.c_int, ip, C.POINTER(ll)]
    L.mdsctk_knn_csc_fetch.argtypes = [vp, ip, dp]
    L.mdsctk_knn_data_rows.argtypes = [vp, dp, ll, C.c_int, dp]
    L.mdsctk_knn_spectral_decomp.argtypes = [vp, C.c_int, ip, ip, dp, C.c_int, C.c_double, C.c_int, dp, dp, dp, dp, ip]
    L.mdsctk_knn_spectral_decomp_ex.argtypes = [vp, C.c_int, ip, ip, dp, C.c_int, C.c_double, C.c_double, C.c_int, dp, dp, dp, dp, ip, dp]
    L.mdsctk_knn_phipsi.argtypes = [vp, fp, ll, C.c_int, dp, dp]
    L.mdsctk_knn_sincos.argtypes = [vp, dp, ll, dp]
    _LIB = L
    return L


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


class DeviceArray:
    """A ctx-owned device array exposed through __cuda_array_interface__ (uint8, 1-D) so that
    torch.as_tensor(arr, device="cuda") can all-gather it over NCCL without a copy."""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class KnnContext:
    """One context per GPU (include/mdsctk_knn.h)."""

    def __init__(self, device=0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.mdsctk_knn_create(C.byref(h), int(device))
        if rc != 0:
            raise KnnError(f"mdsctk_knn_create({device}) -> {rc}: {self._L.mdsctk_knn_last_error(None).decode()}")
        self._h = h
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._L.mdsctk_knn_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc, what):
        if rc != 0:
            raise KnnError(f"{what} -> {rc}: {self._L.mdsctk_knn_last_error(self._h).decode()}")

    def set_option(self, key, value):
        self._ck(self._L.mdsctk_knn_set_option(self._h, key.encode(), int(value)), f"set_option({key})")

    def stats(self):
        s = Stats()
        self._ck(self._L.mdsctk_knn_get_stats(self._h, C.byref(s)), "get_stats")
        return s.as_dict()

    # ---- RMSD path -------------------------------------------------------------------------
    def rms_set_reference(self, xyz, mass):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        mass = np.ascontiguousarray(mass, dtype=np.float32)
        assert xyz.ndim == 3 and xyz.shape[2] == 3 and mass.shape == (xyz.shape[1],)
        self._ck(self._L.mdsctk_knn_rms_set_reference(self._h, _ptr(xyz, C.c_float), xyz.shape[0], xyz.shape[1],
                                                      _ptr(mass, C.c_float)), "rms_set_reference")
        self._n_ref = xyz.shape[0]

    def rms_alloc_reference(self, n_total, n_atoms, mass):
        mass = np.ascontiguousarray(mass, dtype=np.float32)
        self._ck(self._L.mdsctk_knn_rms_alloc_reference(self._h, n_total, n_atoms, _ptr(mass, C.c_float)),
                 "rms_alloc_reference")
        self._n_ref = n_total

    def rms_pack_shard(self, xyz, frame_offset):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self._ck(self._L.mdsctk_knn_rms_pack_shard(self._h, _ptr(xyz, C.c_float), frame_offset, xyz.shape[0]),
                 "rms_pack_shard")

    def _arrays(self, fn, n_rows):
        n = C.c_int(0)
        ptrs = (C.c_void_p * 16)()
        bpf = (C.c_size_t * 16)()
        self._ck(fn(self._h, 16, C.byref(n), ptrs, bpf), "reference_arrays")
        return [DeviceArray(ptrs[i], bpf[i] * n_rows) for i in range(n.value)], [bpf[i] for i in range(n.value)]

    def rms_reference_arrays(self):
        return self._arrays(self._L.mdsctk_knn_rms_reference_arrays, self._n_ref)

    def rms_query(self, k1, fit=None, do_fit=True, fit_range=None, fetch=True):
        """k1 neighbours INCLUDING sorted position 0.  Returns (dist[n_fit,k1] A, idx[n_fit,k1])."""
        if fit is not None:
            fit = np.ascontiguousarray(fit, dtype=np.float32)
            n_fit = fit.shape[0]
        elif fit_range is not None:
            n_fit = fit_range[1]
        else:
            n_fit = self._n_ref
        dist = np.empty((n_fit, k1), dtype=np.float64) if fetch else None
        idx = np.empty((n_fit, k1), dtype=np.int32) if fetch else None
        if fit_range is not None:
            rc = self._L.mdsctk_knn_rms_query_range(self._h, fit_range[0], fit_range[1], k1, int(do_fit),
                                                    _ptr(dist, C.c_double), _ptr(idx, C.c_int))
        else:
            rc = self._L.mdsctk_knn_rms_query(self._h, _ptr(fit, C.c_float), n_fit, k1, int(do_fit),
                                              _ptr(dist, C.c_double), _ptr(idx, C.c_int))
        self._ck(rc, "rms_query")
        return dist, idx

    def rms_rows(self, fit_begin, n_fit, do_fit=True):
        out = np.empty((n_fit, self._n_ref), dtype=np.float64)
        self._ck(self._L.mdsctk_knn_rms_rows(self._h, fit_begin, n_fit, int(do_fit), _ptr(out, C.c_double)), "rms_rows")
        return out

    # ---- vector path -----------------------------------------------------------------------
    def data_set_reference(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        assert rows.ndim == 2
        self._ck(self._L.mdsctk_knn_data_set_reference(self._h, _ptr(rows, C.c_double), rows.shape[0], rows.shape[1]),
                 "data_set_reference")
        self._dn_ref = rows.shape[0]

    def data_alloc_reference(self, n_total, dim):
        self._ck(self._L.mdsctk_knn_data_alloc_reference(self._h, n_total, dim), "data_alloc_reference")
        self._dn_ref = n_total

    def data_upload_shard(self, rows, row_offset):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        self._ck(self._L.mdsctk_knn_data_upload_shard(self._h, _ptr(rows, C.c_double), row_offset, rows.shape[0]),
                 "data_upload_shard")

    def data_reference_arrays(self):
        return self._arrays(self._L.mdsctk_knn_data_reference_arrays, self._dn_ref)

    def data_query(self, k1, fit=None, metric=EUCLIDEAN, fit_range=None, fetch=True):
        if fit is not None:
            fit = np.ascontiguousarray(fit, dtype=np.float64)
            n_fit = fit.shape[0]
        elif fit_range is not None:
            n_fit = fit_range[1]
        else:
            n_fit = self._dn_ref
        dist = np.empty((n_fit, k1), dtype=np.float64) if fetch else None
        idx = np.empty((n_fit, k1), dtype=np.int32) if fetch else None
        if fit_range is not None:
            rc = self._L.mdsctk_knn_data_query_range(self._h, fit_range[0], fit_range[1], k1, int(metric),
                                                     _ptr(dist, C.c_double), _ptr(idx, C.c_int))
        else:
            rc = self._L.mdsctk_knn_data_query(self._h, _ptr(fit, C.c_double), n_fit, k1, int(metric),
                                               _ptr(dist, C.c_double), _ptr(idx, C.c_int))
        self._ck(rc, "data_query")
        return dist, idx

    def data_rows(self, fit=None, metric=EUCLIDEAN):
        """Full distance rows [n_fit, n_reference] (knn_data --sort false)."""
        n_fit = self._dn_ref
        if fit is not None:
            fit = np.ascontiguousarray(fit, dtype=np.float64)
            n_fit = fit.shape[0]
        out = np.empty((n_fit, self._dn_ref), dtype=np.float64)
        self._ck(self._L.mdsctk_knn_data_rows(self._h, _ptr(fit, C.c_double), n_fit, int(metric), _ptr(out, C.c_double)),
                 "data_rows")
        return out

    def spectral_decomp(self, pcol, irow, val, nev, k_sigma=0, sigma=0.0, k_perplexity=0.0, want_sigmas=False):
        """auto_decomp_sparse (k_sigma > 0, -K k_perplexity) / decomp_sparse (k_sigma = 0): (evals[nev] largest first,
        evecs[nev, n], residuals[nev], average sigma, converged pairs[, sigmas[n]])."""
        pcol = np.ascontiguousarray(pcol, dtype=np.int32)
        irow = np.ascontiguousarray(irow, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        n = pcol.size - 1
        ev, vec, res = np.empty(nev), np.empty((nev, n)), np.empty(nev)
        avg, nconv = C.c_double(0.0), C.c_int(0)
        sig = np.empty(n) if want_sigmas else None
        self._ck(self._L.mdsctk_knn_spectral_decomp_ex(self._h, n, _ptr(pcol, C.c_int), _ptr(irow, C.c_int), _ptr(val, C.c_double),
                                                       int(k_sigma), float(sigma), float(k_perplexity), int(nev), _ptr(ev, C.c_double),
                                                       _ptr(vec, C.c_double), _ptr(res, C.c_double), C.byref(avg), C.byref(nconv),
                                                       _ptr(sig, C.c_double)), "spectral_decomp")
        return (ev, vec, res, avg.value, nconv.value) + ((sig,) if want_sigmas else ())

    def phipsi(self, xyz, want_angles=True, want_sincos=True):
        """Backbone torsions (bb_xtc_to_phipsi) and their sin/cos embedding (angles_to_sincos) of N-CA-C frames."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        n, a = xyz.shape[0], xyz.shape[1]
        t = 2 * (a // 3) - 2
        ang = np.empty((n, t), dtype=np.float64) if want_angles else None
        sc = np.empty((n, 2 * t), dtype=np.float64) if want_sincos else None
        self._ck(self._L.mdsctk_knn_phipsi(self._h, _ptr(xyz, C.c_float), n, a, _ptr(ang, C.c_double), _ptr(sc, C.c_double)),
                 "phipsi")
        return ang, sc

    def sincos(self, angles):
        angles = np.ascontiguousarray(angles, dtype=np.float64)
        out = np.empty(angles.size * 2, dtype=np.float64)
        self._ck(self._L.mdsctk_knn_sincos(self._h, _ptr(angles, C.c_double), angles.size, _ptr(out, C.c_double)), "sincos")
        return out

    def csc_build_general(self, idx, dist, k=None, symmetric=False):
        """General CSC matrix of the kNN lists, as make_gesparse [-s] writes it."""
        return self.csc_build_sym(idx, dist, k, _mode=2 if symmetric else 1)

    def csc_build_sym(self, idx, dist, k=None, _mode=0):
        """Symmetric CSC matrix of the kNN lists, as make_sysparse writes it: (pcol[n+1], irow[nnz], val[nnz])."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        dist = np.ascontiguousarray(dist, dtype=np.float64)
        n, maxk = idx.shape
        k = maxk if k is None else k
        pcol = np.empty(n + 1, dtype=np.int32)
        nnz = C.c_longlong(0)
        if _mode == 0:
            rc = self._L.mdsctk_knn_csc_build_sym(self._h, _ptr(idx, C.c_int), _ptr(dist, C.c_double), n, maxk, k,
                                                  _ptr(pcol, C.c_int), C.byref(nnz))
        else:
            rc = self._L.mdsctk_knn_csc_build_general(self._h, _ptr(idx, C.c_int), _ptr(dist, C.c_double), n, maxk, k,
                                                      1 if _mode == 2 else 0, _ptr(pcol, C.c_int), C.byref(nnz))
        self._ck(rc, "csc_build")
        irow = np.empty(nnz.value, dtype=np.int32)
        val = np.empty(nnz.value, dtype=np.float64)
        self._ck(self._L.mdsctk_knn_csc_fetch(self._h, _ptr(irow, C.c_int), _ptr(val, C.c_double)), "csc_fetch")
        return pcol, irow, val

    def debug_fetch_tile(self):
        out = np.empty((128, 9, 48), dtype=np.float32)
        self._ck(self._L.mdsctk_knn_debug_fetch_tile(self._h, _ptr(out, C.c_float)), "debug_fetch_tile")
        return out

    ARRAYS = {"raw": (0, np.float32), "G": (1, np.float32), "cen": (2, np.float64), "sig": (3, np.float32), "Gh": (4, np.float32),
              "G2": (5, np.float32), "gres": (6, np.float32), "planes": (7, np.float32), "hi": (8, np.float32), "lo": (9, np.float32),
              "bh": (10, np.uint16), "bm": (11, np.uint16), "fh": (12, np.float16), "fl": (13, np.float16)}

    def debug_fetch_array(self, name):
        """One packed array of the RMSD reference set as a flat numpy array (tests of the pack kernel)."""
        which, dt = self.ARRAYS[name]
        nb = C.c_size_t(0)
        self._ck(self._L.mdsctk_knn_debug_fetch_array(self._h, which, None, 0, C.byref(nb)), "debug_fetch_array")
        out = np.empty(nb.value // np.dtype(dt).itemsize, dtype=dt)
        self._ck(self._L.mdsctk_knn_debug_fetch_array(self._h, which, out.ctypes.data_as(C.c_void_p), nb.value, C.byref(nb)),
                 "debug_fetch_array")
        return out

    def timer_start(self):
        self._ck(self._L.mdsctk_knn_timer_start(self._h), "timer_start")

    def timer_stop(self):
        ms = C.c_double(0.0)
        self._ck(self._L.mdsctk_knn_timer_stop(self._h, C.byref(ms)), "timer_stop")
        return ms.value

    def fetch(self, n_fit, k1):
        dist = np.empty((n_fit, k1), dtype=np.float64)
        idx = np.empty((n_fit, k1), dtype=np.int32)
        self._ck(self._L.mdsctk_knn_fetch(self._h, _ptr(dist, C.c_double), _ptr(idx, C.c_int)), "fetch")
        return dist, idx


def _clamp_k(k, n_ref):
    # knn_rms.cpp:224-225 / knn_data.cpp:186-188: k = min(k, n_ref - 1), k1 = k + 1
    return max(0, min(int(k), n_ref - 1))


def knn_rms(ref_xyz, mass, k, fit_xyz=None, nofit=False, device=0, rms_kernel=None, ctx=None):
    """What `knn_rms -k K [-n true] -p top -r ref [-f fit]` writes: (distances[n_fit,k], indices[n_fit,k])."""
    own = ctx is None
    ctx = ctx or KnnContext(device)
    try:
        if rms_kernel is not None:
            ctx.set_option("rms_kernel", rms_kernel)
        ctx.rms_set_reference(ref_xyz, mass)
        k = _clamp_k(k, ref_xyz.shape[0])
        dist, idx = ctx.rms_query(k + 1, fit=fit_xyz, do_fit=not nofit)
        return np.ascontiguousarray(dist[:, 1:]), np.ascontiguousarray(idx[:, 1:])
    finally:
        if own:
            ctx.close()


def knn_data(ref_rows, k, fit_rows=None, correlation=False, device=0, ctx=None):
    """What `knn_data -k K -v D [-c] -r ref [-f fit]` writes."""
    own = ctx is None
    ctx = ctx or KnnContext(device)
    try:
        ctx.data_set_reference(ref_rows)
        k = _clamp_k(k, ref_rows.shape[0])
        dist, idx = ctx.data_query(k + 1, fit=fit_rows, metric=CORRELATION if correlation else EUCLIDEAN)
        return np.ascontiguousarray(dist[:, 1:]), np.ascontiguousarray(idx[:, 1:])
    finally:
        if own:
            ctx.close()
