"""GPU parity tests of the tcgen05 (tensor-core) RMSD sweep: raw TMEM accumulators against a
numpy contraction, then the same end-to-end parity bar as the FP32 kernel."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TC_3X, TC_1X, TC_BF, TC_F3, TC_F2, TC_F1 = 1, 2, 3, 4, 5, 6


@pytest.fixture(scope="module")
def ctx():
    import mdsctk_b200
    c = mdsctk_b200.KnnContext(0)
    yield c
    c.close()


def packed_planes(xyz, mass):
    w = mass.astype(np.float64) / mass.astype(np.float64).sum()
    x = xyz.astype(np.float64)
    c = (x * w[None, :, None]).sum(axis=1, keepdims=True)
    return ((x - c) * np.sqrt(w)[None, :, None]).astype(np.float32).astype(np.float64)   # [n, A, 3]


@pytest.mark.parametrize("kern,rtol", [(TC_3X, 2e-6), (TC_1X, 2e-3), (TC_BF, 3e-5), (TC_F3, 2e-6), (TC_F2, 5e-4), (TC_F1, 1e-3)])
def test_tmem_accumulators_match_numpy(ctx, trpcage, kern, rtol):
    xyz, mass = trpcage
    xyz = xyz[:300]
    ctx.set_option("rms_kernel", kern)
    ctx.set_option("debug_tile", 1)
    try:
        ctx.rms_set_reference(xyz, mass)
        ctx.rms_query(11)
        tile = ctx.debug_fetch_tile()            # [128 q][3a+b][48 j]
    finally:
        ctx.set_option("debug_tile", 0)
        ctx.set_option("rms_kernel", 0)
    P = packed_planes(xyz, mass)
    want = np.einsum("qna,jnb->qabj", P[:128], P[:48]).reshape(128, 9, 48)
    scale = np.einsum("qna,jnb->qabj", np.abs(P[:128]), np.abs(P[:48])).reshape(128, 9, 48)
    err = np.abs(tile - want) / scale.max()
    assert err.max() < rtol, f"max scaled error {err.max()}"


@pytest.mark.parametrize("atoms,version,wide", [(300, 2, 1), (300, 2, 0), (300, 1, 1), (256, 2, 1), (290, 2, 1), (290, 2, 0), (33, 2, 1),
                                                (33, 2, 0), (20, 2, 1), (5, 2, 1), (304, 2, 1), (128, 2, 1), (64, 2, 1), (80, 2, 1), (288, 2, 1), (280, 2, 0),
                                                (224, 2, 1), (200, 2, 1), (240, 2, 1), (176, 2, 1)])
def test_tmem_accumulators_1xfp16_resident_tile(ctx, atoms, version, wide):
    """The default sweep (rms_tc2.cu) keeps the fit tile in shared memory and, beyond 256 atoms, its trailing k-steps in
    TMEM (tcgen05.mma with the A operand in tensor memory): raw accumulators against numpy on the fp16-rounded operands.
    wide: 64-atom ring stages of 128-byte rows (SWIZZLE_128B reference operand against the SWIZZLE_64B fit tile), the
    default where three fit; 0: 32-atom stages."""
    from mdsctk_b200 import synth
    xyz = synth.traj_frames(400, atoms, 2, 9)
    mass = (12.0 + np.arange(atoms) % 3).astype(np.float32)
    ctx.set_option("rms_kernel", TC_F1)
    ctx.set_option("sweep_version", version)
    ctx.set_option("rms_wide_stages", wide)
    ctx.set_option("debug_tile", 1)
    try:
        ctx.rms_set_reference(xyz, mass)
        ctx.rms_query(11)
        assert ctx.stats()["sweep_version"] == version
        tile = ctx.debug_fetch_tile()
    finally:
        ctx.set_option("debug_tile", 0)
        ctx.set_option("rms_wide_stages", 1)
        ctx.set_option("sweep_version", 2)
        ctx.set_option("rms_kernel", 0)
    P = packed_planes(xyz, mass)
    H = (P.astype(np.float32) * np.float32(64)).astype(np.float16).astype(np.float64) / 64.0      # what the tensor cores see
    want = np.einsum("qna,jnb->qabj", H[:128], H[:48]).reshape(128, 9, 48)
    scale = np.einsum("qna,jnb->qabj", np.abs(H[:128]), np.abs(H[:48])).reshape(128, 9, 48)
    err = np.abs(tile - want) / scale.max()
    assert err.max() < 5e-6, f"max scaled error {err.max()}"          # fp32 accumulation only: the operands are exact


@pytest.mark.parametrize("kern", [TC_3X, TC_1X, TC_BF, TC_F3, TC_F2, TC_F1])
@pytest.mark.parametrize("k", [10, 100])
def test_trpcage_knn_rms_tc(ctx, trpcage, kern, k):
    import mdsctk_b200
    xyz, mass = trpcage
    g = np.load(os.path.join(GOLDEN, f"trpcage_rms_k{k}.npz"))
    dist, idx = mdsctk_b200.knn_rms(xyz, mass, k, ctx=ctx, rms_kernel=kern)
    st = ctx.stats()
    ctx.set_option("rms_kernel", 0)
    assert st["rms_kernel"] == kern
    assert np.array_equal(idx, g["idx_f64"])
    assert (np.abs(dist - g["dist_f64"]) <= 1e-9 * g["dist_f64"]).all()
    assert (np.abs(dist - g["dist_ref"]) <= 1e-4 * g["dist_ref"]).all()
    print("tc kernel", kern, "k", k, {n: st[n] for n in ("fallback_rows", "rescored_max", "max_filter_err", "max_filter_spread",
                                                         "cert_eps", "cert_gres", "k_keep")})
    if kern != TC_1X:                   # the coarse TF32 filter may fall back to exact rows; the result is the same
        assert st["fallback_rows"] <= 2
        assert 0.5 * st["max_filter_spread"] < st["cert_eps"]
    if kern in (TC_F2, TC_F1):          # operand-rounding term of the certificate: residual norms from pack.cu
        assert 0.0 < st["cert_gres"] < (2e-6 if kern == TC_F2 else 2e-3)


@pytest.mark.parametrize("kern", [TC_3X, TC_1X, TC_BF, TC_F3, TC_F2, TC_F1])
def test_synthetic_300_atoms_tc(ctx, kern):
    import mdsctk_b200
    from mdsctk_b200 import synth
    from oracle import binding as ob
    n = 5000
    xyz = synth.traj_frames(n, 300, 16)
    mass = synth.traj_masses(300)
    dist, idx = mdsctk_b200.knn_rms(xyz, mass, 32, ctx=ctx, rms_kernel=kern)
    st = ctx.stats()
    ctx.set_option("rms_kernel", 0)
    d, i = ob.knn_rms(xyz, mass, 32, fit=xyz[:256], mode=1)
    assert np.array_equal(idx[:256], i)
    assert (np.abs(dist[:256] - d) <= 1e-9 * d).all()
    assert (np.diff(dist, axis=1) >= 0).all() and (idx != np.arange(n)[:, None]).all()
    # both kernels must agree on every row (the FP64 stage decides, the sweep only filters)
    dist0, idx0 = mdsctk_b200.knn_rms(xyz, mass, 32, ctx=ctx, rms_kernel=0)
    assert np.array_equal(idx, idx0) and np.array_equal(dist, dist0)
    print("tc kernel", kern, {k: st[k] for k in ("ms_sweep", "fallback_rows", "rescored_max", "max_filter_err", "max_filter_spread",
                                                 "cert_eps", "cert_gres", "k_keep", "lists_per_row")})
    if kern != TC_1X:
        assert st["fallback_rows"] <= 2


def test_rounding_residuals_bound_the_filter_error(ctx):
    """1xFP16: |sqrt(key) - d| <= g_q + g_r up to accumulation noise -- the metric argument the certificate rests on.
    Checked on the raw accumulators of one tile: the RMSD between the rounded structures (numpy, FP64) must equal the
    kernel's key, and the residual norms numpy computes must match the stat the certificate used."""
    import mdsctk_b200
    from mdsctk_b200 import synth
    n = 2000
    xyz = synth.traj_frames(n, 300, 4)
    mass = synth.traj_masses(300)
    P = packed_planes(xyz, mass)                               # float32-rounded operands, as pack.cu forms them
    hi = (P * 64.0).astype(np.float16).astype(np.float64) / 64.0
    w = mass.astype(np.float64) / mass.astype(np.float64).sum()
    x = xyz.astype(np.float64)
    T = (x - (x * w[None, :, None]).sum(axis=1, keepdims=True)) * np.sqrt(w)[None, :, None]
    g1 = np.sqrt(((T - hi) ** 2).sum(axis=(1, 2)))
    ctx.set_option("debug_tile", 1)
    try:
        dist, idx = mdsctk_b200.knn_rms(xyz, mass, 32, ctx=ctx, rms_kernel=TC_F1)
        st = ctx.stats()
        tile = ctx.debug_fetch_tile()                          # [128 q][3a+b][48 j]: hi x hi contraction
    finally:
        ctx.set_option("debug_tile", 0)
        ctx.set_option("rms_kernel", 0)
    assert abs(st["cert_gres"] - g1.max()) <= 1e-6 * g1.max() + 1e-9
    want = np.einsum("qna,jnb->qabj", hi[:128], hi[:48]).reshape(128, 9, 48)
    assert np.abs(tile - want).max() <= 3e-6 * np.abs(want).max()      # products of fp16 values are exact; fp32 accumulation
    dist0, idx0 = mdsctk_b200.knn_rms(xyz, mass, 32, ctx=ctx, rms_kernel=0)
    assert np.array_equal(idx, idx0) and np.array_equal(dist, dist0)
    assert st["fallback_rows"] == 0


def test_default_kernel_and_fp16_overflow_substitution(trpcage):
    """The default is the 1xFP16 sweep; coordinates that would overflow fp16 (64*sqrt(G) > 3e4) make the default give
    way to 3xTF32, while an explicitly requested fp16 kernel is refused.  Same neighbours either way (RMSD scales)."""
    import mdsctk_b200
    xyz, mass = trpcage
    xyz = xyz[:300]
    with mdsctk_b200.KnnContext(0) as c:
        dist, idx = mdsctk_b200.knn_rms(xyz, mass, 10, ctx=c)
        assert c.stats()["rms_kernel"] == TC_F1
        big = (xyz * np.float32(1000.0)).astype(np.float32)          # G ~ 5e5 nm^2
        dist_b, idx_b = mdsctk_b200.knn_rms(big, mass, 10, ctx=c)
        assert c.stats()["rms_kernel"] == TC_3X
        assert np.array_equal(idx_b, idx) and np.allclose(dist_b, dist * 1000.0, rtol=1e-5)
        with pytest.raises(mdsctk_b200.KnnError):
            mdsctk_b200.knn_rms(big, mass, 10, ctx=c, rms_kernel=TC_F1)


def test_out_of_sample_start_tile_guess(ctx):
    """Fit rows handed over as host frames (the -f path) that happen to be reference frames must give the bytes of
    the in-sample range query: the singular-value start-tile guess only changes the order of the sweep."""
    from mdsctk_b200 import synth
    n = 6000
    xyz = synth.traj_frames(n, 300, 8)
    mass = synth.traj_masses(300)
    ctx.set_option("rms_kernel", TC_F1)
    try:
        ctx.rms_set_reference(xyz, mass)
        d0, i0 = ctx.rms_query(33, fit_range=(2500, 700))
        d1, i1 = ctx.rms_query(33, fit=np.ascontiguousarray(xyz[2500:3200]))
        assert ctx.stats()["fallback_rows"] == 0
    finally:
        ctx.set_option("rms_kernel", 0)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    assert (i0[:, 0] == np.arange(2500, 3200)).all() and (d0[:, 0] < 1e-5).all()       # rank 0 is the frame itself (sqrt of FP64 noise)


def test_ragged_and_out_of_sample_tc(ctx, trpcage):
    import mdsctk_b200
    from oracle import binding as ob
    xyz, mass = trpcage
    for n in (7, 49, 129, 200):
        dist, idx = mdsctk_b200.knn_rms(xyz[:n], mass, 5, ctx=ctx, rms_kernel=TC_BF)
        d, i = ob.knn_rms(xyz[:n], mass, min(5, n - 1), mode=1)
        assert np.array_equal(idx, i), n
    g = np.load(os.path.join(GOLDEN, "trpcage_rms_oos_k10.npz"))
    fit, ref = xyz[::10], np.delete(xyz, np.arange(0, 1000, 10), axis=0)
    dist, idx = mdsctk_b200.knn_rms(ref, mass, 10, fit_xyz=fit, ctx=ctx, rms_kernel=TC_3X)
    assert np.array_equal(idx, g["idx_f64"])
    g = np.load(os.path.join(GOLDEN, "trpcage_rms_nofit_k10.npz"))
    dist, idx = mdsctk_b200.knn_rms(xyz, mass, 10, nofit=True, ctx=ctx, rms_kernel=TC_3X)
    ctx.set_option("rms_kernel", 0)
    assert np.array_equal(idx, g["idx_f64"])
