"""auto_decomp_sparse: the oracle's affinity stage against an independent numpy
restatement of auto_decomp_sparse.cpp:150-198 and its dense eigen-solve against numpy.linalg.eigh on CPU; the GPU
adjacency / affinity / thick-restart Lanczos (spectral.cu) against the oracle; the tool end to end."""
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


def knn_graph(n, k, seed=0, dim=3):
    from oracle import binding as ob
    rng = np.random.default_rng(seed)
    # three blobs: a clustered graph has a clear spectral gap
    pts = np.concatenate([rng.normal(size=(n // 3, dim)) + 6 * c for c in range(3)] + [rng.normal(size=(n - 3 * (n // 3), dim))])
    d, i = ob.knn_data(pts, k)
    return ob.make_sysparse(i, d)


def numpy_affinity(pcol, irow, val, k_a):
    n = len(pcol) - 1
    lists = [[] for _ in range(n)]
    for x in range(n):
        for y in range(pcol[x], pcol[x + 1]):
            lists[x].append(val[y])
            lists[irow[y]].append(val[y])
    sigma = np.array([sum(sorted(l[:k_a])) / k_a for l in lists])
    m = np.array(val, dtype=np.float64)
    col = np.repeat(np.arange(n), np.diff(pcol))
    m = np.exp(-(m * m) / (2.0 * sigma[col] * sigma[irow]))
    deg = np.zeros(n)
    np.add.at(deg, col, m)
    np.add.at(deg, irow, m)
    dinv = 1.0 / np.sqrt(deg)
    return m * dinv[irow] * dinv[col], sigma.mean()


def dense(pcol, irow, m):
    n = len(pcol) - 1
    a = np.zeros((n, n))
    col = np.repeat(np.arange(n), np.diff(pcol))
    a[col, irow] = m
    a[irow, col] = m
    return a


def align(v, ref):
    """flip the sign of every row of v to match ref (eigenvector signs are arbitrary)"""
    s = np.sign(np.einsum("ij,ij->i", v, ref))
    s[s == 0] = 1
    return v * s[:, None]


def test_oracle_affinity_and_eigs_against_numpy():
    from oracle import binding as ob
    pcol, irow, val = knn_graph(150, 8, seed=1)
    m, avg = ob.affinity(pcol, irow, val, 5)
    m2, avg2 = numpy_affinity(pcol, irow, val, 5)
    assert np.allclose(m, m2, rtol=1e-13, atol=0) and abs(avg - avg2) < 1e-13
    ev, vec, res = ob.sym_eigs_largest(pcol, irow, m, 6)
    w, v = np.linalg.eigh(dense(pcol, irow, m))
    assert np.allclose(ev, w[::-1][:6], rtol=1e-12, atol=1e-13)
    assert abs(ev[0] - 1.0) < 1e-12                           # D^-1/2 W D^-1/2 has the trivial eigenvalue 1
    assert res.max() < 1e-10
    top = v[:, ::-1][:, :7].T
    for e in range(6):      # eigenvectors are unique (up to sign) only where the eigenvalue is isolated
        if min(abs(w[::-1][e] - w[::-1][e + 1]), abs(w[::-1][e] - w[::-1][e - 1]) if e else 1.0) > 1e-6:
            assert abs(abs(vec[e] @ top[e]) - 1.0) < 1e-8
        assert np.linalg.norm(dense(pcol, irow, m) @ vec[e] - ev[e] * vec[e]) < 1e-10


def test_auto_decomp_sparse_cli_contract(tmp_path):
    def run(*args):
        p = subprocess.run([os.path.join(BIN, "auto_decomp_sparse"), *args], capture_output=True, text=True, cwd=tmp_path)
        return p.returncode, p.stdout
    rc, out = run("-h")
    assert rc == 1 and "usage: auto_decomp_sparse [options]" in out
    for opt in ("--k-sigma", "--k-perplexity", "--nevals", "--ssm-file", "--evals-file", "--evecs-file", "--residuals-file",
                "(=distances.ssm)", "(=eigenvalues.dat)", "(=eigenvectors.dat)", "(=residuals.dat)"):   # auto_decomp_sparse.cpp:58-67
        assert opt in out
    rc, out = run()
    assert rc == 255 and "ERROR: --k-sigma not supplied." in out and "ERROR: --nevals not supplied." in out   # :79-88
    rc, out = run("-k", "5", "-n", "3")
    assert rc == 3 and "k-sigma =        5" in out and "nevals =         3" in out and "ssm-file =       distances.ssm" in out


@pytest.mark.gpu
def test_gpu_spectral_matches_oracle():
    import mdsctk_b200
    from oracle import binding as ob
    with mdsctk_b200.KnnContext(0) as ctx:
        for n, k, k_a, nev, seed in ((300, 10, 5, 4, 2), (600, 12, 12, 10, 3), (90, 6, 3, 2, 4)):
            pcol, irow, val = knn_graph(n, k, seed=seed)
            m, avg = ob.affinity(pcol, irow, val, k_a)
            want_ev, want_vec, _ = ob.sym_eigs_largest(pcol, irow, m, nev)
            ev, vec, res, gavg, nconv = ctx.spectral_decomp(pcol, irow, val, nev, k_sigma=k_a)
            assert abs(gavg - avg) < 1e-12 * avg
            assert nconv == nev
            assert np.allclose(ev, want_ev, rtol=1e-10, atol=1e-12), (n, ev, want_ev)
            assert res.max() < 1e-9
            a = dense(pcol, irow, m)
            # compare invariant subspaces where eigenvalues are (nearly) degenerate, vectors elsewhere
            gaps = np.abs(np.diff(np.concatenate([want_ev, [np.linalg.eigvalsh(a)[::-1][nev]]])))
            for e in range(nev):
                if min(gaps[e], gaps[e - 1] if e else 1.0) > 1e-6:
                    assert abs(abs(vec[e] @ want_vec[e]) - 1.0) < 1e-7, (n, e)
                assert abs(np.linalg.norm(vec[e]) - 1.0) < 1e-12
                assert np.linalg.norm(a @ vec[e] - ev[e] * vec[e]) < 1e-9
            # decomp_sparse mode: the matrix as given (already normalised affinities)
            ev2, vec2, res2, _, nconv2 = ctx.spectral_decomp(pcol, irow, m, nev, k_sigma=0)
            assert np.allclose(ev2, want_ev, rtol=1e-10, atol=1e-12) and nconv2 == nev


@pytest.mark.gpu
def test_auto_decomp_sparse_tool_end_to_end(tmp_path):
    """kNN lists (trp-cage golden) -> make_sysparse -> auto_decomp_sparse: text files as the reference writes them
    (auto_decomp_sparse.cpp:207-236: eigenvalues largest first, one eigenvector per line, residuals)."""
    from oracle import binding as ob
    g = np.load(os.path.join(GOLDEN, "trpcage_rms_k10.npz"))
    idx, dist = g["idx_f64"].astype(np.int32), g["dist_f64"]
    idx.tofile(tmp_path / "indices.dat")
    dist.tofile(tmp_path / "distances.dat")
    p = subprocess.run([os.path.join(BIN, "make_sysparse"), "-k", "10"], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 0, p.stdout
    p = subprocess.run([os.path.join(BIN, "auto_decomp_sparse"), "-k", "5", "-n", "5"], capture_output=True, text=True, cwd=tmp_path)
    assert p.returncode == 0, p.stdout
    assert "Average sigma: " in p.stdout and "Number of converged eigenvalues/vectors found: 5" in p.stdout
    assert "Maximum residual: " in p.stdout
    pcol, irow, val = ob.make_sysparse(idx, dist)
    m, avg = ob.affinity(pcol, irow, val, 5)
    want_ev, want_vec, _ = ob.sym_eigs_largest(pcol, irow, m, 5)
    ev = np.loadtxt(tmp_path / "eigenvalues.dat")
    vec = np.loadtxt(tmp_path / "eigenvectors.dat")
    res = np.loadtxt(tmp_path / "residuals.dat")
    assert ev.shape == (5,) and vec.shape == (5, 1000) and res.shape == (5,)
    assert np.allclose(ev, want_ev, rtol=2e-6)                # the reference prints 6 significant digits
    assert (res < 1e-8).all()
    assert ("Average sigma: %g" % avg) in p.stdout


def first_k_sorted(pcol, irow, val, k_a):
    """sorted_A of auto_decomp_sparse.cpp:153-165: per frame the first k_a values in CSC traversal order, sorted."""
    n = len(pcol) - 1
    lists = [[] for _ in range(n)]
    for x in range(n):
        for y in range(pcol[x], pcol[x + 1]):
            lists[x].append(val[y])
            lists[irow[y]].append(val[y])
    return [sorted(l[:k_a]) for l in lists]


def perplexity(dist, sigma):
    p = np.exp(-np.asarray(dist) ** 2 / (2.0 * sigma * sigma))
    p /= p.sum()
    return float(np.exp(-(p * np.log(p)).sum()))


def test_oracle_entropic_sigmas_have_the_requested_perplexity():
    """entropic_affinity_sigmas (mdsctk.cpp:498-565): the Gaussian over a frame's sorted distances with the returned
    sigma has perplexity K -- a known-answer property of the restatement, warm-start chain included."""
    from oracle import binding as ob
    rng = np.random.default_rng(1)
    for k, K in ((32, 10.0), (12, 4.5), (64, 30.0)):
        a = np.sort(rng.random((500, k)) * rng.uniform(0.5, 4.0, (500, 1)) + 0.1, axis=1)
        s = ob.entropic_sigmas(a, K)
        got = np.array([perplexity(a[i], s[i]) for i in range(500)])
        assert np.abs(got - K).max() < 1e-7 * K


@pytest.mark.gpu
def test_gpu_entropic_affinities_match_the_oracle(tmp_path):
    """auto_decomp_sparse -K: per-frame entropic sigmas on the GPU (independent frames, midpoint start) against the
    oracle's sequential warm-start chain, the perplexity property itself, and the tool's option."""
    import mdsctk_b200
    from oracle import binding as ob
    n, k, k_a, K, nev = 900, 16, 16, 6.0, 4
    # ONE elongated blob: a connected kNN graph, so the leading eigenvalue 1 is simple (the three-blob graph of the other
    # tests is disconnected at k = 16, and a single-vector Lanczos -- like ARPACK -- finds a threefold eigenvalue only by luck)
    pts = np.random.default_rng(7).normal(size=(n, 3)) * np.array([3.0, 1.0, 1.0])
    d_, i_ = ob.knn_data(pts, k)
    pcol, irow, val = ob.make_sysparse(i_, d_)
    rows = first_k_sorted(pcol, irow, val, k_a)
    assert all(len(r) == k_a for r in rows)                     # every frame has k_a entries: the reference's precondition
    want = ob.entropic_sigmas(np.array(rows), K)
    with mdsctk_b200.KnnContext(0) as ctx:
        ev, vec, res, avg, nconv, sig = ctx.spectral_decomp(pcol, irow, val, nev, k_sigma=k_a, k_perplexity=K, want_sigmas=True)
        ev0, _, _, avg0, _ = ctx.spectral_decomp(pcol, irow, val, nev, k_sigma=k_a)
    assert np.abs(sig - want).max() < 1e-8 * want.max()
    got = np.array([perplexity(rows[i], sig[i]) for i in range(n)])
    assert np.abs(got - K).max() < 1e-7 * K
    assert abs(avg - want.mean()) < 1e-8 * avg and abs(avg - avg0) > 1e-3 * avg0      # not the mean-distance sigmas
    assert nconv == nev and res.max() < 1e-9 and ev[0] > ev[-1] > 0
    # the same matrix by numpy with the oracle's sigmas
    col = np.repeat(np.arange(n), np.diff(pcol))
    m = np.exp(-(val * val) / (2.0 * want[col] * want[irow]))
    deg = np.zeros(n); np.add.at(deg, col, m); np.add.at(deg, irow, m)
    a = dense(pcol, irow, m / np.sqrt(deg[col] * deg[irow]))
    assert np.allclose(ev, np.linalg.eigvalsh(a)[::-1][:nev], rtol=1e-8)
    # the tool
    with open(tmp_path / "distances.ssm", "wb") as f:
        np.array([n], dtype=np.int32).tofile(f); pcol.astype(np.int32).tofile(f); irow.astype(np.int32).tofile(f); val.tofile(f)
    p = subprocess.run([os.path.join(BIN, "auto_decomp_sparse"), "-k", str(k_a), "-K", str(K), "-n", str(nev)], capture_output=True,
                       text=True, cwd=tmp_path)
    assert p.returncode == 0, p.stdout
    assert "k-perplexity =   6" in p.stdout and ("Average sigma: %g" % avg) in p.stdout
    assert np.allclose(np.loadtxt(tmp_path / "eigenvalues.dat"), ev, rtol=2e-6)
