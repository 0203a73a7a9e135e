"""knn_data_sparse: the oracle's merge distance against its dense form on CPU, the
CLI contract, and on the GPU the tool's files against the oracle, bit for bit (sorted, out of sample, full rows)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "mdsctk_b200", "bin")


@pytest.fixture(scope="module", autouse=True)
def built():
    from mdsctk_b200 import build
    build.build_all()


def random_sparse(rng, n, ndim, max_nnz):
    out = []
    for _ in range(n):
        m = int(rng.integers(0, max_nnz + 1))
        idx = np.sort(rng.choice(ndim, m, replace=False)).astype(np.int32)
        out.append((idx, rng.normal(size=m)))
    return out


def write_sparse(vectors, index_path, data_path):
    with open(index_path, "wb") as fi, open(data_path, "wb") as fd:      # contact_profile.cpp's layout: int n; int idx[n] / double val[n]
        for idx, val in vectors:
            fi.write(np.int32(len(idx)).tobytes())
            fi.write(np.asarray(idx, dtype=np.int32).tobytes())
            fd.write(np.asarray(val, dtype=np.float64).tobytes())


def test_oracle_sparse_distance_equals_dense_in_union_order():
    from oracle import binding as ob
    rng = np.random.default_rng(4)
    vs = random_sparse(rng, 60, 500, 12)
    d, i = ob.knn_data_sparse(vs, 7)
    dims = np.unique(np.concatenate([v[0] for v in vs]))
    dense = np.zeros((60, dims.size))
    for r, (idx, val) in enumerate(vs):
        dense[r, np.searchsorted(dims, idx)] = val
    d2, i2 = ob.knn_data(dense, 7)
    assert np.array_equal(d, d2) and np.array_equal(i, i2)          # bit-identical: what the GPU tool relies on


def test_knn_data_sparse_cli_contract(tmp_path):
    p = subprocess.run([os.path.join(BIN, "knn_data_sparse"), "-h"], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 1 and "usage: knn_data_sparse [options]" in p.stdout
    for opt in ("--reference-index-file", "--reference-data-file", "--fit-index-file", "--fit-data-file", "(=reference.svi)",
                "(=reference.svd)"):                                     # knn_data_sparse.cpp:72-83
        assert opt in p.stdout
    p = subprocess.run([os.path.join(BIN, "knn_data_sparse")], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 255 and "ERROR: --knn not supplied." in p.stdout
    p = subprocess.run([os.path.join(BIN, "knn_data_sparse"), "-k", "3"], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 3 and "reference-index-file = reference.svi" in p.stdout and "fit-file =             reference.svd" in p.stdout


@pytest.mark.gpu
def test_knn_data_sparse_tool_matches_oracle(tmp_path):
    from oracle import binding as ob
    rng = np.random.default_rng(5)
    ref = random_sparse(rng, 700, 4000, 25)
    fit = random_sparse(rng, 90, 4200, 25)
    write_sparse(ref, tmp_path / "reference.svi", tmp_path / "reference.svd")
    write_sparse(fit, tmp_path / "fit.svi", tmp_path / "fit.svd")
    tool = os.path.join(BIN, "knn_data_sparse")
    p = subprocess.run([tool, "-k", "9"], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 0, p.stdout
    d, i = ob.knn_data_sparse(ref, 9)
    assert np.array_equal(np.fromfile(tmp_path / "indices.dat", dtype=np.int32).reshape(700, 9), i)
    assert np.array_equal(np.fromfile(tmp_path / "distances.dat", dtype=np.float64).reshape(700, 9), d)
    p = subprocess.run([tool, "-k", "9", "-F", "fit.svi", "-f", "fit.svd"], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 0, p.stdout
    d, i = ob.knn_data_sparse(ref, 9, fit=fit)
    assert np.array_equal(np.fromfile(tmp_path / "indices.dat", dtype=np.int32).reshape(90, 9), i)
    assert np.array_equal(np.fromfile(tmp_path / "distances.dat", dtype=np.float64).reshape(90, 9), d)
    p = subprocess.run([tool, "-s", "false", "-F", "fit.svi", "-f", "fit.svd"], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 0, p.stdout
    rows = np.fromfile(tmp_path / "distances.dat", dtype=np.float64).reshape(90, 700)
    L = ob.lib()
    import ctypes as C
    L.oracle_euclidean_distance_sparse.restype = C.c_double
    L.oracle_euclidean_distance_sparse.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int),
                                                   C.POINTER(C.c_double)]
    for f in (0, 41, 89):
        for r in (0, 350, 699):
            ri, rv = np.ascontiguousarray(ref[r][0]), np.ascontiguousarray(ref[r][1])
            fi, fv = np.ascontiguousarray(fit[f][0]), np.ascontiguousarray(fit[f][1])
            want = L.oracle_euclidean_distance_sparse(len(ri), ri.ctypes.data_as(C.POINTER(C.c_int)), ob._d(rv), len(fi),
                                                      fi.ctypes.data_as(C.POINTER(C.c_int)), ob._d(fv))
            assert rows[f, r] == want
