"""Parity at BASELINE.json's full single-GPU sizes through size-independent properties (sortedness, no self
neighbour, symmetry of mutual neighbours, agreement of the two tensor-core splits) plus the oracle on a row sample.
C3: 100 000 frames x 300 atoms, k=32.  knn_data: C5's shape at 200 000 rows x 512 dims, k=64."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def mutual_symmetry(dist, idx, rows):
    """for sampled rows i and their neighbours j that list i back: d(i,j) must equal d(j,i)"""
    worst, seen = 0.0, 0
    for i in rows:
        for pos, j in enumerate(idx[i]):
            back = np.nonzero(idx[j] == i)[0]
            if back.size:
                worst = max(worst, abs(dist[j, back[0]] - dist[i, pos]) / dist[i, pos])
                seen += 1
    return worst, seen


def test_c3_full_size_properties_and_oracle_sample():
    import mdsctk_b200
    from mdsctk_b200 import synth
    from oracle import binding as ob
    n, k = 100_000, 32
    xyz = synth.traj_frames(n, 300, 16, 20260117)
    mass = synth.traj_masses(300)
    with mdsctk_b200.KnnContext(0) as ctx:
        dist, idx = mdsctk_b200.knn_rms(xyz, mass, k, ctx=ctx)                      # default kernel (1xFP16)
        st = ctx.stats()
        assert st["rms_kernel"] == mdsctk_b200.RMS_TC_1XFP16 and st["fallback_rows"] <= 8
        assert dist.shape == (n, k) and (np.diff(dist, axis=1) >= 0).all()
        assert (idx != np.arange(n)[:, None]).all() and (idx >= 0).all() and (idx < n).all()
        assert all(len(set(r)) == k for r in idx[::997])
        rng = np.random.default_rng(0)
        worst, seen = mutual_symmetry(dist, idx, rng.integers(0, n, 300))
        assert seen > 1000 and worst < 1e-9                                         # FP64 re-score on both sides
        # the reference's float chain (oracle mode 0) and the FP64 Kabsch (mode 1) on a row sample against all frames
        rows = np.concatenate([np.arange(0, 64), np.arange(50_000, 50_032), np.arange(n - 32, n)])
        d1, i1 = ob.knn_rms(xyz, mass, k, fit=xyz[rows], mode=1)
        assert np.array_equal(idx[rows], i1) and (np.abs(dist[rows] - d1) <= 1e-9 * d1).all()
        d0, i0 = ob.knn_rms(xyz, mass, k, fit=xyz[rows[:48]], mode=0)
        assert (np.abs(dist[rows[:48]] - d0) <= 1e-4 * d0).all()
        # the BF16 split must give the same bytes (the FP64 stage decides, the contraction only filters)
        dist3, idx3 = mdsctk_b200.knn_rms(xyz, mass, k, ctx=ctx, rms_kernel=mdsctk_b200.RMS_TC_3XBF16)
        assert np.array_equal(idx3, idx) and np.array_equal(dist3, dist)
        # and so must the one-MMA and the full three-MMA fp16 sweeps, whichever of them is the default
        for kern in (mdsctk_b200.RMS_TC_1XFP16, mdsctk_b200.RMS_TC_3XFP16):
            dist1, idx1 = mdsctk_b200.knn_rms(xyz, mass, k, ctx=ctx, rms_kernel=kern)
            st1 = ctx.stats()
            print("kernel", kern, {n_: st1[n_] for n_ in ("ms_sweep", "ms_rescore", "ms_fallback", "fallback_rows", "rescored_max",
                                                        "max_filter_err", "cert_eps", "cert_gres", "k_keep")})
            assert np.array_equal(idx1, idx) and np.array_equal(dist1, dist)
            assert st1["fallback_rows"] <= 8


def test_knn_data_c5_shape_properties_and_oracle_sample():
    import mdsctk_b200
    from mdsctk_b200 import synth
    from oracle import binding as ob
    n, dim, k = 200_000, 512, 64
    rows = synth.phipsi_rows(n, dim, 64)
    with mdsctk_b200.KnnContext(0) as ctx:
        dist, idx = mdsctk_b200.knn_data(rows, k, ctx=ctx)
        st = ctx.stats()
        assert st["lists_per_row"] >= 1 and st["k_keep"] > k and st["fallback_rows"] < 50     # the tensor path ran
        assert (np.diff(dist, axis=1) >= 0).all() and (idx != np.arange(n)[:, None]).all()
        rng = np.random.default_rng(1)
        worst, seen = mutual_symmetry(dist, idx, rng.integers(0, n, 200))
        assert seen > 1000 and worst == 0.0                                         # bit-identical both ways
        sample = np.concatenate([np.arange(0, 24), np.arange(n - 24, n)])
        d, i = ob.knn_data(rows, k, fit=rows[sample])
        assert np.array_equal(idx[sample], i) and np.array_equal(dist[sample], d)
