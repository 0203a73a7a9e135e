"""GPU parity tests of the tensor-core knn_data path (data_tc.cu): the tcgen05 contraction only
filters, the FP64 re-score reproduces the CPU tool's arithmetic, so results must be BIT-IDENTICAL
to the oracle / golden files -- on the reference's own example inputs, on wide rows, out of
sample, and on ragged sizes."""
import os

import numpy as np
import pytest

from conftest import DATA, GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import mdsctk_b200
    c = mdsctk_b200.KnnContext(0)
    c.set_option("data_kernel", 1)
    yield c
    c.close()


@pytest.mark.parametrize("name,dim,k", [("rings", 2, 10), ("rings", 2, 20), ("swissroll", 3, 10), ("swissroll", 3, 12)])
def test_examples_bit_exact_tc(ctx, name, dim, k):
    import mdsctk_b200
    pts = np.fromfile(os.path.join(DATA, f"{name}.pts"), dtype=np.float64).reshape(-1, dim)
    g = np.load(os.path.join(GOLDEN, f"{name}_data_k{k}.npz"))
    dist, idx = mdsctk_b200.knn_data(pts, k, ctx=ctx)
    st = ctx.stats()
    assert st["lists_per_row"] >= 1 and st["k_keep"] > k      # the tensor path ran (it sets both)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])


def test_wide_rows_tc_matches_oracle(ctx):
    import mdsctk_b200
    from mdsctk_b200 import synth
    from oracle import binding as ob
    rows = synth.phipsi_rows(20000, 512, 16)
    dist, idx = mdsctk_b200.knn_data(rows, 64, ctx=ctx)
    st = ctx.stats()
    d, i = ob.knn_data(rows, 64, fit=rows[:256])
    assert np.array_equal(idx[:256], i) and np.array_equal(dist[:256], d)
    d, i = ob.knn_data(rows, 64, fit=rows[19744:])
    assert np.array_equal(idx[19744:], i) and np.array_equal(dist[19744:], d)
    assert (np.diff(dist, axis=1) >= 0).all() and (idx != np.arange(20000)[:, None]).all()
    assert st["fallback_rows"] < 200
    assert 0.5 * st["max_filter_spread"] <= st["cert_eps"]
    print("data tc", {k: st[k] for k in ("ms_pack", "ms_sweep", "ms_rescore", "ms_fallback", "fallback_rows", "max_filter_err",
                                         "max_filter_spread", "cert_eps", "k_keep", "lists_per_row", "rescored_max")})
    # the exact FP64 sweep gives the same bytes
    ctx.set_option("data_kernel", 0)
    try:
        dist0, idx0 = mdsctk_b200.knn_data(rows[:4000], 64, ctx=ctx)
    finally:
        ctx.set_option("data_kernel", 1)
    dist1, idx1 = mdsctk_b200.knn_data(rows[:4000], 64, ctx=ctx)
    assert np.array_equal(idx0, idx1) and np.array_equal(dist0, dist1)


def test_out_of_sample_and_ragged_tc(ctx):
    import mdsctk_b200
    from mdsctk_b200 import synth
    from oracle import binding as ob
    rows = synth.phipsi_rows(3000, 100, 8)            # dim not a multiple of 32, n not a multiple of 256
    fit = rows[::7] + 1e-3
    dist, idx = mdsctk_b200.knn_data(rows, 10, fit_rows=fit, ctx=ctx)
    d, i = ob.knn_data(rows, 10, fit=fit)
    assert np.array_equal(idx, i) and np.array_equal(dist, d)
    for n in (5, 129, 257, 1000):
        dist, idx = mdsctk_b200.knn_data(rows[:n], 4, ctx=ctx)
        d, i = ob.knn_data(rows[:n], min(4, n - 1))
        assert np.array_equal(idx, i) and np.array_equal(dist, d), n
    # duplicated rows: exact ties are ordered (distance, index); whole plateaus force the exact fallback
    dup = np.concatenate([rows[:500], rows[:500]])
    dist, idx = mdsctk_b200.knn_data(dup, 6, ctx=ctx)
    d, i = ob.knn_data(dup, 6)
    assert np.array_equal(idx, i) and np.array_equal(dist, d)


def test_large_values_and_scale_tc(ctx):
    import mdsctk_b200
    from oracle import binding as ob
    rng = np.random.default_rng(7)
    rows = rng.normal(size=(2000, 40)) * 1.0e4 + 3.0e5          # far from the origin: |x|^2 >> d^2
    dist, idx = mdsctk_b200.knn_data(rows, 8, ctx=ctx)
    d, i = ob.knn_data(rows, 8)
    assert np.array_equal(idx, i) and np.array_equal(dist, d)
    rows = rng.normal(size=(1500, 16)) * 1.0e-7
    dist, idx = mdsctk_b200.knn_data(rows, 8, ctx=ctx)
    d, i = ob.knn_data(rows, 8)
    assert np.array_equal(idx, i) and np.array_equal(dist, d)


def test_one_part_filter_gives_the_same_bytes():
    """data_kernel 2: the fp16 hi parts alone feed the filter (one MMA per k-step); the re-score's certificate bounds
    the rounding by the triangle inequality, so the files stay bit-identical to the three-MMA filter and the oracle."""
    import mdsctk_b200
    from mdsctk_b200 import synth
    from oracle import binding as ob
    rows = synth.phipsi_rows(20000, 512, 16)
    with mdsctk_b200.KnnContext(0) as c:
        c.set_option("data_kernel", 2)
        dist, idx = mdsctk_b200.knn_data(rows, 64, ctx=c)
        st = c.stats()
        print("data one-part", {k: st[k] for k in ("ms_sweep", "ms_rescore", "ms_fallback", "fallback_rows", "max_filter_err",
                                                   "cert_eps", "cert_gres", "k_keep", "rescored_max")})
        assert st["cert_gres"] > 0.0 and st["fallback_rows"] < 200
        d, i = ob.knn_data(rows, 64, fit=rows[:256])
        assert np.array_equal(idx[:256], i) and np.array_equal(dist[:256], d)
        c.set_option("data_kernel", 1)
        dist1, idx1 = mdsctk_b200.knn_data(rows, 64, ctx=c)
        assert np.array_equal(idx, idx1) and np.array_equal(dist, dist1)
        # the reference's small examples and an out-of-sample query
        pts = np.fromfile(os.path.join(DATA, "swissroll.pts"), dtype=np.float64).reshape(-1, 3)
        g = np.load(os.path.join(GOLDEN, "swissroll_data_k10.npz"))
        c.set_option("data_kernel", 2)
        dist, idx = mdsctk_b200.knn_data(pts, 10, ctx=c)
        assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
        dist, idx = mdsctk_b200.knn_data(rows[:3000], 20, fit_rows=rows[5000:5400], ctx=c)
        d, i = ob.knn_data(rows[:3000], 20, fit=rows[5000:5400])
        assert np.array_equal(idx, i) and np.array_equal(dist, d)


@pytest.mark.parametrize("kernel", [1, 2])
def test_correlation_metric_on_the_tensor_path_is_bit_exact(ctx, kernel):
    """knn_data -c (knn_data.cpp:104-106, correlation_distance mdsctk.cpp:337-360): the Euclidean tensor filter on the
    standardised rows + exact FP64 re-score with the reference's own arithmetic must reproduce the oracle's bytes, in
    sample and out of sample, and agree with the exact FP64 sweep; constant rows send the query to the exact sweep."""
    import mdsctk_b200
    from mdsctk_b200 import synth
    from oracle import binding as ob
    rng = np.random.default_rng(4)
    rows = synth.phipsi_rows(20000, 128, 16) * rng.uniform(0.5, 3.0, (20000, 1)) + rng.normal(0, 2.0, (20000, 1))   # per-row scale and offset
    ctx.set_option("data_kernel", kernel)
    try:
        dist, idx = mdsctk_b200.knn_data(rows, 24, correlation=True, ctx=ctx)
        st = ctx.stats()
        assert st["lists_per_row"] >= 1 and st["k_keep"] > 25 and st["fallback_rows"] < 400      # the tensor path ran
        sample = np.concatenate([np.arange(0, 128), np.arange(19900, 20000)])
        d, i = ob.knn_data(rows, 24, fit=rows[sample], metric=1)
        assert np.array_equal(idx[sample], i) and np.array_equal(dist[sample], d)
        fit = synth.phipsi_rows(3000, 128, 16, seed=77) * 1.7 - 0.3
        dist_o, idx_o = mdsctk_b200.knn_data(rows, 24, fit_rows=fit, correlation=True, ctx=ctx)
        d, i = ob.knn_data(rows, 24, fit=fit[:200], metric=1)
        assert np.array_equal(idx_o[:200], i) and np.array_equal(dist_o[:200], d)
        ctx.set_option("data_kernel", 0)
        dist0, idx0 = mdsctk_b200.knn_data(rows[:6000], 24, correlation=True, ctx=ctx)
        ctx.set_option("data_kernel", kernel)
        dist1, idx1 = mdsctk_b200.knn_data(rows[:6000], 24, correlation=True, ctx=ctx)
        assert ctx.stats()["lists_per_row"] >= 1
        assert np.array_equal(idx0, idx1) and np.array_equal(dist0, dist1)
        # the Euclidean metric right after it on the same context: the packed operands are rebuilt for the metric
        dist_e, idx_e = mdsctk_b200.knn_data(rows[:6000], 24, ctx=ctx)
        d, i = ob.knn_data(rows[:6000], 24, fit=rows[:100])
        assert np.array_equal(idx_e[:100], i) and np.array_equal(dist_e[:100], d)
        # a constant row has zero spread: correlation_distance divides by it -- exact sweep, whatever it makes of it
        bad = rows[:6000].copy()
        bad[17] = 3.0
        dist_b, idx_b = mdsctk_b200.knn_data(bad, 24, correlation=True, ctx=ctx)
        ctx.set_option("data_kernel", 0)
        dist_b0, idx_b0 = mdsctk_b200.knn_data(bad, 24, correlation=True, ctx=ctx)
        assert np.array_equal(idx_b, idx_b0) and np.array_equal(dist_b, dist_b0, equal_nan=True)
    finally:
        ctx.set_option("data_kernel", 1)
