"""CPU tests that pin the oracle (oracle/) -- the checker the GPU path is graded against.

The reference has no tests or golden outputs, so the oracle is
pinned by (i) committed golden vectors generated from it on the reference's
example inputs, (ii) an independent numpy SVD Kabsch, and (iii) the invariance
properties listed in SURVEY.md Appendix B.
"""
import os

import numpy as np
import pytest

from conftest import DATA, GOLDEN, load_pts
from oracle import binding as ob


def kabsch_numpy(a, b, m):
    w = m.astype(np.float64)
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    a = a - (w[:, None] * a).sum(0) / w.sum()
    b = b - (w[:, None] * b).sum(0) / w.sum()
    S = (a * w[:, None]).T @ b
    U, s, Vt = np.linalg.svd(S)
    s[2] *= np.sign(np.linalg.det(U @ Vt))
    msd = ((w[:, None] * a * a).sum() + (w[:, None] * b * b).sum() - 2 * s.sum()) / w.sum()
    return np.sqrt(max(0.0, msd))


def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_xtc_decode_matches_golden(trpcage):
    xyz, mass = trpcage
    g = np.load(os.path.join(GOLDEN, "trpcage_decode.npz"))
    assert xyz.shape == (1000, 60, 3)
    grid = np.round(xyz.astype(np.float64) * 1000).astype(np.int64)
    assert np.array_equal(grid.sum(axis=(1, 2)), g["colsum"])
    assert np.array_equal(np.array([grid.sum(), (grid * grid).sum()]), g["total"])
    assert np.array_equal(xyz[0], g["frame0"]) and np.array_equal(xyz[999], g["frame999"])
    assert np.array_equal(mass, g["mass"])


def test_xtc_decode_is_physical(trpcage):
    xyz, _ = trpcage
    d = np.linalg.norm(xyz[:, 1:] - xyz[:, :-1], axis=2)
    # backbone N-CA, CA-C, C-N bond lengths (nm)
    assert abs(d[:, 0::3].mean() - 0.147) < 0.003
    assert abs(d[:, 1::3].mean() - 0.154) < 0.003
    assert abs(d[:, 2::3].mean() - 0.134) < 0.003
    assert d.max() < 0.2
    # xtc precision 1000: coordinates sit on the 0.001 nm grid
    assert np.abs(xyz * 1000 - np.round(xyz * 1000)).max() < 1e-3


def test_masses_from_pdb(trpcage):
    _, mass = trpcage
    assert mass.shape == (60,)
    assert np.allclose(mass[0::3], 14.0067) and np.allclose(mass[1::3], 12.0107) and np.allclose(mass[2::3], 12.0107)


def test_rmsd_f64_matches_numpy_svd(trpcage):
    xyz, mass = trpcage
    rng = np.random.default_rng(1)
    for i, j in rng.integers(0, 1000, size=(50, 2)):
        assert ob.rmsd_f64(xyz[i], xyz[j], mass) == pytest.approx(kabsch_numpy(xyz[i], xyz[j], mass), rel=1e-10, abs=1e-9)


def test_float_chain_close_to_f64(trpcage):
    xyz, mass = trpcage
    rng = np.random.default_rng(2)
    for i, j in rng.integers(0, 1000, size=(100, 2)):
        if i == j:
            continue
        r0, _ = ob.fit_rmsdev(ob.reset_x(xyz[i], mass), ob.reset_x(xyz[j], mass), mass)
        r1 = ob.rmsd_f64(xyz[i], xyz[j], mass)
        assert r0 == pytest.approx(r1, rel=2e-5)


def test_rigid_motion_invariance(trpcage):
    xyz, mass = trpcage
    rng = np.random.default_rng(3)
    a, b = xyz[10].astype(np.float64), xyz[700].astype(np.float64)
    base = ob.rmsd_f64(a, b, mass)
    for _ in range(5):
        Ra, Rb = random_rotation(rng), random_rotation(rng)
        a2 = (a @ Ra.T + rng.uniform(-1, 1, 3)).astype(np.float32)
        b2 = (b @ Rb.T + rng.uniform(-1, 1, 3)).astype(np.float32)
        assert ob.rmsd_f64(a2, b2, mass) == pytest.approx(base, rel=1e-5)
        # a frame against a rigidly moved copy of itself: zero
        assert ob.rmsd_f64(a.astype(np.float32), a2, mass) < 2e-6
        r0, _ = ob.fit_rmsdev(ob.reset_x(a.astype(np.float32), mass), ob.reset_x(a2, mass), mass)
        assert r0 < 1e-5
    # symmetry
    assert ob.rmsd_f64(a, b, mass) == pytest.approx(ob.rmsd_f64(b, a, mass), rel=1e-12)


def test_mirror_image_is_not_superposable(trpcage):
    xyz, mass = trpcage
    a = xyz[0]
    mirrored = a * np.array([1, 1, -1], dtype=np.float32)
    assert ob.rmsd_f64(a, mirrored, mass) > 0.05
    r0, _ = ob.fit_rmsdev(ob.reset_x(a, mass), ob.reset_x(mirrored, mass), mass)
    assert r0 == pytest.approx(ob.rmsd_f64(a, mirrored, mass), rel=1e-4)


def test_nofit_is_plain_rmsd(trpcage):
    xyz, mass = trpcage
    a, b = ob.reset_x(xyz[3], mass), ob.reset_x(xyz[4], mass)
    w = mass.astype(np.float64)
    want = np.sqrt((w[:, None] * (a.astype(np.float64) - b) ** 2).sum() / w.sum())
    got, _ = ob.fit_rmsdev(a, b, mass, dofit=False)
    assert got == pytest.approx(want, rel=1e-6)
    assert ob.rmsd_f64(xyz[3], xyz[4], mass, dofit=False) == pytest.approx(want, rel=1e-6)


def test_three_atom_known_answer():
    # equilateral triangle vs the same triangle stretched along x: analytic minimum
    a = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    m = np.ones(3, dtype=np.float32)
    assert ob.rmsd_f64(a, a + 5, m) < 1e-7
    b = a.copy()
    b[1, 0] = 2.0
    assert ob.rmsd_f64(a, b, m) == pytest.approx(kabsch_numpy(a, b, m), rel=1e-12)


@pytest.mark.parametrize("k", [10, 100])
def test_knn_rms_golden(trpcage, k):
    xyz, mass = trpcage
    g = np.load(os.path.join(GOLDEN, f"trpcage_rms_k{k}.npz"))
    d1, i1 = ob.knn_rms(xyz, mass, k, mode=1)
    assert np.array_equal(i1, g["idx_f64"]) and np.allclose(d1, g["dist_f64"], rtol=1e-12, atol=0)
    # rows ascending, self removed
    assert (np.diff(d1, axis=1) >= 0).all()
    assert (i1 != np.arange(1000)[:, None]).all()
    # reference float chain vs FP64: same neighbours at k=10, distances within 1e-4
    assert np.abs(g["dist_ref"] - d1).max() / d1.min() < 1e-4
    if k == 10:
        assert np.array_equal(g["idx_ref"], i1)
    else:
        assert (np.sort(g["idx_ref"], axis=1) == np.sort(i1, axis=1)).all(axis=1).mean() > 0.99


def test_knn_rms_float_chain_golden_subset(trpcage):
    xyz, mass = trpcage
    g = np.load(os.path.join(GOLDEN, "trpcage_rms_k10.npz"))
    d0, i0 = ob.knn_rms(xyz, mass, 10, fit=xyz[:64], mode=0)
    assert np.array_equal(i0, g["idx_ref"][:64]) and np.array_equal(d0, g["dist_ref"][:64])


@pytest.mark.parametrize("name,dim,k", [("rings", 2, 10), ("rings", 2, 20), ("swissroll", 3, 10), ("swissroll", 3, 12)])
def test_knn_data_golden_bit_exact(name, dim, k):
    pts = load_pts(f"{name}.pts", dim)
    g = np.load(os.path.join(GOLDEN, f"{name}_data_k{k}.npz"))
    d, i = ob.knn_data(pts, k)
    assert np.array_equal(i, g["idx"]) and np.array_equal(d, g["dist"])
    # numpy restatement of mdsctk.cpp:330-335 with the same summation order
    acc = np.zeros((pts.shape[0], pts.shape[0]))
    for x in range(dim):
        diff = pts[:, None, x] - pts[None, :, x]
        acc = acc + diff * diff
    full = np.sqrt(acc)
    order = np.argsort(full, axis=1, kind="stable")[:, 1:k + 1]
    assert np.array_equal(order, i)
    assert np.array_equal(np.take_along_axis(full, order, axis=1), d)


def test_correlation_distance_matches_pearson():
    rng = np.random.default_rng(5)
    a, b = rng.normal(size=64), rng.normal(size=64)
    r = np.corrcoef(a, b)[0, 1]
    assert ob.correlation_distance(a, b) == pytest.approx(np.sqrt((1 - r) / 2), rel=1e-12)
    assert ob.correlation_distance(a, a) < 1e-7


def test_partial_row_is_dropped():
    # knn_data.cpp:146-151 drops a trailing partial row; load_pts mirrors that
    raw = np.fromfile(os.path.join(DATA, "rings.pts"), dtype=np.float64)
    assert load_pts("rings.pts", 3).shape == (raw.size // 3, 3)


def test_qcp_c0_identity():
    """det K = |S|_F^4 - 4 |cof S|_F^2 for the 4x4 key matrix K of a 3x3 cross-covariance S: the form
    mdsctk_b200/csrc/qcp.cuh uses for the constant coefficient of the QCP characteristic polynomial."""
    rng = np.random.default_rng(7)
    for _ in range(500):
        S = rng.standard_normal((3, 3)) * rng.uniform(0.01, 3.0, (3, 1))
        (sxx, sxy, sxz), (syx, syy, syz), (szx, szy, szz) = S
        K = np.array([[sxx + syy + szz, syz - szy, szx - sxz, sxy - syx],
                      [syz - szy, sxx - syy - szz, sxy + syx, szx + sxz],
                      [szx - sxz, sxy + syx, -sxx + syy - szz, syz + szy],
                      [sxy - syx, szx + sxz, syz + szy, -sxx - syy + szz]])
        F = (S ** 2).sum()
        cof = np.array([[np.linalg.det(np.delete(np.delete(S, i, 0), j, 1)) for j in range(3)] for i in range(3)])
        assert abs(np.linalg.det(K) - (F * F - 4.0 * (cof ** 2).sum())) <= 1e-12 * F * F


def test_rounded_structure_triangle_bound():
    """The certificate of the 1xFP16 sweep: min-RMSD between the fp16-rounded structures is within
    g_q + g_r (the rounding residual norms) of the true min-RMSD.  Checked with the oracle's FP64 Kabsch."""
    from mdsctk_b200 import synth
    xyz = synth.traj_frames(600, 100, 3, 5).astype(np.float64)
    w = np.full(100, 1.0 / 100)
    T = (xyz - (xyz * w[None, :, None]).sum(axis=1, keepdims=True)) * np.sqrt(w)[None, :, None]
    R = (T * 64.0).astype(np.float32).astype(np.float16).astype(np.float64) / 64.0
    g = np.sqrt(((T - R) ** 2).sum(axis=(1, 2)))

    def min_rmsd(a, b):                       # rotation only (the rounded structures are not re-centred)
        u, s, vt = np.linalg.svd(a.T @ b)
        s[-1] *= np.sign(np.linalg.det(u @ vt))
        return np.sqrt(max((a ** 2).sum() + (b ** 2).sum() - 2.0 * s.sum(), 0.0))

    rng = np.random.default_rng(1)
    for q, r in rng.integers(0, 600, (300, 2)):
        assert abs(min_rmsd(T[q], T[r]) - min_rmsd(R[q], R[r])) <= g[q] + g[r] + 1e-12
