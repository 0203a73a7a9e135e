"""make_sysparse (the consumer of the kNN files): oracle known answers and CLI
contract on CPU; the GPU builder (csc.cu) against the oracle, bit for bit, on kNN graphs, graphs with
duplicate entries, hub columns, k < maxk, and end to end through the tool on the trp-cage golden lists."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

BIN = os.path.join(ROOT, "mdsctk_b200", "bin")


@pytest.fixture(scope="module", autouse=True)
def built():
    from mdsctk_b200 import build
    build.build_all()


def reference_semantics(idx, dist, k):
    """make_sysparse.cpp:245-277 as a dict: later insertions overwrite (Db::put), rows in order."""
    n = idx.shape[0]
    db = {}
    for i in range(n):
        for x in range(k):
            j = int(idx[i, x])
            if i < j:
                db[(i, j)] = dist[i, x]
            if j < i:
                db[(j, i)] = dist[i, x]
    keys = sorted(db)                                 # compare_edge order, mdsctk.cpp:567-577
    pcol = np.zeros(n + 1, dtype=np.int32)
    for (f, _t) in keys:
        if 0 <= f < n:
            pcol[f + 1] += 1
    keys = [kk for kk in keys if 0 <= kk[0] < n]
    return np.cumsum(pcol, dtype=np.int32), np.array([t for _f, t in keys], dtype=np.int32), \
        np.array([db[kk] for kk in keys], dtype=np.float64)


def reference_semantics_general(idx, dist, k, symmetric):
    """make_gesparse.cpp:246-275 as a dict: put always; with -s put the transposed entry only while it is absent."""
    n = idx.shape[0]
    db = {}
    for i in range(n):
        for x in range(k):
            j = int(idx[i, x])
            db[(i, j)] = dist[i, x]
            if symmetric and (j, i) not in db:
                db[(j, i)] = dist[i, x]
    keys = [kk for kk in sorted(db) if 0 <= kk[0] < n]
    pcol = np.zeros(n + 1, dtype=np.int32)
    for (f, _t) in keys:
        pcol[f + 1] += 1
    return np.cumsum(pcol, dtype=np.int32), np.array([t for _f, t in keys], dtype=np.int32), \
        np.array([db[kk] for kk in keys], dtype=np.float64)


def random_lists(rng, n, maxk, dup=False, hub=False):
    idx = np.empty((n, maxk), dtype=np.int32)
    for i in range(n):
        if dup:
            idx[i] = rng.integers(0, n, maxk)         # duplicates and self loops allowed
        else:
            idx[i] = rng.choice(n, maxk, replace=False)
    if hub:
        idx[:, 0] = 3                                 # every row lists frame 3: a column with ~n entries
    dist = rng.random((n, maxk))
    return idx, dist


def test_oracle_matches_reference_semantics():
    from oracle import binding as ob
    rng = np.random.default_rng(1)
    idx = np.array([[1, 2], [0, 2], [1, 0]], dtype=np.int32)
    d = np.array([[1.0, 2.0], [1.5, 3.0], [3.5, 2.5]])
    pcol, irow, val = ob.make_sysparse(idx, d)
    assert pcol.tolist() == [0, 2, 3, 3] and irow.tolist() == [1, 2, 2] and val.tolist() == [1.5, 2.5, 3.5]
    for n, maxk, k, dup in ((50, 6, 6, False), (80, 9, 4, True), (33, 5, 0, False), (200, 12, 12, True)):
        idx, dist = random_lists(rng, n, maxk, dup=dup)
        got = ob.make_sysparse(idx, dist, k)
        want = reference_semantics(idx, dist, k)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)


def test_gesparse_oracle_matches_reference_semantics():
    from oracle import binding as ob
    rng = np.random.default_rng(3)
    for n, maxk, k, dup in ((40, 5, 5, False), (90, 8, 3, True), (150, 10, 10, True)):
        idx, dist = random_lists(rng, n, maxk, dup=dup)
        for sym in (False, True):
            got = ob.make_gesparse(idx, dist, k, symmetric=sym)
            want = reference_semantics_general(idx, dist, k, sym)
            for g, w in zip(got, want):
                assert np.array_equal(g, w), (n, maxk, k, dup, sym)


def test_make_sysparse_cli_contract(tmp_path):
    def run(*args):
        p = subprocess.run([os.path.join(BIN, "make_sysparse"), *args], capture_output=True, text=True, cwd=tmp_path)
        return p.returncode, p.stdout
    rc, out = run("-h")
    assert rc == 1 and "usage: make_sysparse [options]" in out           # make_sysparse.cpp:82-86
    for opt in ("--knn", "--output-knn", "--index-file", "--distance-file", "--output-file"):   # :67-73
        assert opt in out
    rc, out = run()
    assert rc == 255 and "ERROR: --knn not supplied." in out              # :87-91,103-105
    rc, out = run("-k", "5", "-n", "7")
    assert rc == 255 and "ERROR: Output k (7) is not less than the input k (5)." in out   # :97-101
    rc, out = run("-k", "5")
    assert rc == 255 and "Could not open file: distances.dat" in out      # :140-146
    assert "knn =            5" in out and "output-knn =     5" in out and "output-file =    distances.ssm" in out
    p = subprocess.run([os.path.join(BIN, "make_gesparse"), "-h"], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 1 and "usage: make_gesparse [options]" in p.stdout and "--symmetric" in p.stdout   # make_gesparse.cpp:66-76
    p = subprocess.run([os.path.join(BIN, "make_gesparse"), "-k", "4"], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 255 and "output-file =    distances.gsm" in p.stdout


@pytest.mark.gpu
def test_gpu_builder_matches_oracle():
    import mdsctk_b200
    from oracle import binding as ob
    rng = np.random.default_rng(2)
    with mdsctk_b200.KnnContext(0) as ctx:
        cases = [(1, 1, 1, False, False), (2, 1, 1, True, False), (300, 10, 10, False, False), (300, 10, 3, False, False),
                 (1000, 8, 8, True, False), (5000, 16, 16, False, True), (2500, 7, 5, True, True), (40000, 32, 32, False, False)]
        for n, maxk, k, dup, hub in cases:
            if dup or maxk > n:
                idx, dist = random_lists(rng, n, min(maxk, max(n, 1)), dup=True, hub=hub)
            else:
                idx, dist = random_lists(rng, n, maxk, hub=hub)
            k = min(k, idx.shape[1])
            got = ctx.csc_build_sym(idx, dist, k)
            want = ob.make_sysparse(idx, dist, k)
            for g, w, name in zip(got, want, ("pcol", "irow", "val")):
                assert np.array_equal(g, w), (n, maxk, k, dup, hub, name)
            for sym in (False, True):
                got = ctx.csc_build_general(idx, dist, k, symmetric=sym)
                want = ob.make_gesparse(idx, dist, k, symmetric=sym)
                for g, w, name in zip(got, want, ("pcol", "irow", "val")):
                    assert np.array_equal(g, w), ("general", sym, n, maxk, k, dup, hub, name)
        # k = 0: empty matrix
        pcol, irow, val = ctx.csc_build_sym(idx, dist, 0)
        assert not pcol.any() and irow.size == 0 and val.size == 0


@pytest.mark.gpu
def test_make_sysparse_tool_end_to_end(tmp_path):
    """kNN files exactly as knn_rms writes them (golden trp-cage lists) -> distances.ssm, byte for byte
    what the oracle's CSC gives, and readable the way CSC_matrix does (mdsctk.cpp:44-59)."""
    from oracle import binding as ob
    g = np.load(os.path.join(GOLDEN, "trpcage_rms_k100.npz"))
    idx, dist = g["idx_f64"].astype(np.int32), g["dist_f64"]
    idx.tofile(tmp_path / "indices.dat")
    dist.tofile(tmp_path / "distances.dat")
    for k in (100, 12):
        p = subprocess.run([os.path.join(BIN, "make_sysparse"), "-k", "100", "-n", str(k)], capture_output=True, text=True,
                           cwd=tmp_path)
        assert p.returncode == 0, p.stdout
        assert "Creating sparse matrix database..." in p.stdout and "Converting database to sparse matrix..." in p.stdout
        raw = (tmp_path / "distances.ssm").read_bytes()
        n = int(np.frombuffer(raw, dtype=np.int32, count=1)[0])
        pcol = np.frombuffer(raw, dtype=np.int32, count=n + 1, offset=4)
        nnz = int(pcol[-1])
        irow = np.frombuffer(raw, dtype=np.int32, count=nnz, offset=4 + 4 * (n + 1))
        val = np.frombuffer(raw, dtype=np.float64, count=nnz, offset=4 + 4 * (n + 1) + 4 * nnz)
        assert len(raw) == 4 + 4 * (n + 1) + 12 * nnz and n == idx.shape[0]
        want = ob.make_sysparse(idx, dist, k)
        assert np.array_equal(pcol, want[0]) and np.array_equal(irow, want[1]) and np.array_equal(val, want[2])
        # strictly upper triangular, rows ascending within a column
        for c in (0, 1, n // 2, n - 2):
            r = irow[pcol[c]:pcol[c + 1]]
            assert (r > c).all() and (np.diff(r) > 0).all()
        for flag, sym in (((), False), (("-s",), True)):
            p = subprocess.run([os.path.join(BIN, "make_gesparse"), "-k", "100", "-n", str(k), *flag], capture_output=True,
                               text=True, cwd=tmp_path)
            assert p.returncode == 0, p.stdout
            raw = (tmp_path / "distances.gsm").read_bytes()
            n = int(np.frombuffer(raw, dtype=np.int32, count=1)[0])
            pcol = np.frombuffer(raw, dtype=np.int32, count=n + 1, offset=4)
            nnz = int(pcol[-1])
            irow = np.frombuffer(raw, dtype=np.int32, count=nnz, offset=4 + 4 * (n + 1))
            val = np.frombuffer(raw, dtype=np.float64, count=nnz, offset=4 + 4 * (n + 1) + 4 * nnz)
            want = ob.make_gesparse(idx, dist, k, symmetric=sym)
            assert np.array_equal(pcol, want[0]) and np.array_equal(irow, want[1]) and np.array_equal(val, want[2])
