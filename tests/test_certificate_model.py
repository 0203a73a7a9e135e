"""CPU model of the 1xFP16 sweep + certificate (mdsctk_b200/csrc/rms_rescore.cu) in numpy.

The GPU kernels cannot run here; what CAN be checked without a GPU is the mathematics the index-exactness rests on:
  * von Neumann pre-bound:  RMSD^2(x, y) >= sum_i (sigma_i(x) - sigma_i(y))^2   (never rejects a true neighbour)
  * rounded-structure bound: |d~ - d| <= g_q + g_r with d~ the min-RMSD of the fp16-rounded structures
  * the certificate rule: whenever  min_i [a_i - max(d_i - g, 0)^2] + (d_k + g)^2 + 2 eps < a_next  holds for keys
    a = d~^2 + bias + noise (|noise| <= eps), the k1 nearest among the re-scored candidates ARE the k1 nearest overall.
The FP64 Kabsch of the CPU checker is the ground truth."""
import numpy as np

from mdsctk_b200 import synth


def lam_max(S):
    """largest eigenvalue of the 4x4 key matrix of 3x3 cross-covariances S[..., 3, 3] (proper rotations only)"""
    sxx, sxy, sxz, syx, syy, syz, szx, szy, szz = [S[..., i, j] for i in range(3) for j in range(3)]
    K = np.empty(S.shape[:-2] + (4, 4))
    K[..., 0, 0] = sxx + syy + szz; K[..., 0, 1] = syz - szy; K[..., 0, 2] = szx - sxz; K[..., 0, 3] = sxy - syx
    K[..., 1, 1] = sxx - syy - szz; K[..., 1, 2] = sxy + syx; K[..., 1, 3] = szx + sxz
    K[..., 2, 2] = -sxx + syy - szz; K[..., 2, 3] = syz + szy
    K[..., 3, 3] = -sxx - syy + szz
    for i in range(4):
        for j in range(i):
            K[..., i, j] = K[..., j, i]
    return np.linalg.eigvalsh(K)[..., -1]


def setup(n=1500, atoms=100, basins=3, seed=11):
    x = synth.traj_frames(n, atoms, basins, seed).astype(np.float64)
    w = 1.0 / atoms
    T = (x - x.mean(axis=1, keepdims=True)) * np.sqrt(w)               # true weighted, centred structures
    H = (T * 64.0).astype(np.float32).astype(np.float16).astype(np.float64) / 64.0   # what the tensor cores contract
    return T, H


def msd_rows(A, q):
    """exact min-RMSD^2 of structure q against all structures of A (rotation only, as the sweep's QCP)"""
    S = np.einsum("ai,raj->rij", A[q], A)
    G = (A ** 2).sum(axis=(1, 2))
    return np.maximum(G[q] + G - 2.0 * lam_max(S), 0.0)


def test_von_neumann_prebound_never_exceeds_the_distance():
    T, _ = setup()
    sig = np.linalg.svd(T, compute_uv=False)                           # [n, 3] descending
    for q in range(0, T.shape[0], 97):
        d2 = msd_rows(T, q)
        lb = ((sig[q][None, :] - sig) ** 2).sum(axis=1)
        assert (lb <= d2 + 1e-12).all()
    # and it separates the basins of the generator: every frame of another basin is further away in singular-value
    # space than any frame of the row's own basin (frames 0..499 / 1000..1499 of 3 x 500)
    lb = ((sig[0][None, :] - sig) ** 2).sum(axis=1)
    assert lb[1000:].min() > lb[:500].max()


def test_rounded_distance_within_residual_norms():
    T, H = setup()
    g = np.sqrt(((T - H) ** 2).sum(axis=(1, 2)))
    for q in range(0, T.shape[0], 131):
        d, dt = np.sqrt(msd_rows(T, q)), np.sqrt(msd_rows(H, q))
        assert (np.abs(d - dt) <= g[q] + g + 1e-12).all()


def certificate(keys, d_exact, order, m, k1, g, eps):
    """the rule of rms_rescore_kernel after re-scoring the m candidates with the smallest keys"""
    idx = order[:m]
    a, d = keys[idx], d_exact[idx]
    dk = np.sort(d)[k1 - 1]
    a_next = keys[order[m]] if m < len(order) else np.inf
    b_up = (a - np.maximum(d - g, 0.0) ** 2).min()
    brackets_ok = (a - (d + g) ** 2).max() - b_up <= 2.0 * eps           # model check
    return brackets_ok and b_up + (dk + g) ** 2 + 2.0 * eps < a_next, set(idx[np.argsort(d, kind="stable")[:k1]])


def test_certificate_never_certifies_a_wrong_neighbour_set():
    T, H = setup()
    n = T.shape[0]
    g_all = np.sqrt(((T - H) ** 2).sum(axis=(1, 2)))
    rng = np.random.default_rng(3)
    certified = 0
    for q in range(0, n, 53):
        d = np.sqrt(msd_rows(T, q))
        e0 = 0.5 * ((T[q] ** 2).sum() + (T ** 2).sum(axis=(1, 2)))
        eps = 2.0e-6 * e0.max()
        bias = 1.3 * eps                                                 # row-common, unknown to the rule
        keys = msd_rows(H, q) + bias + rng.uniform(-eps, eps, n)         # what the sweep hands over
        order = np.argsort(keys, kind="stable")
        g = g_all[q] + g_all.max()
        for k1 in (9, 33):
            truth = set(np.argsort(d, kind="stable")[:k1])
            for m in (k1, 64, 128, 256):
                ok, got = certificate(keys, d, order, m, k1, g, eps)
                if ok:
                    certified += 1
                    assert got == truth, (q, k1, m)
            ok, _ = certificate(keys, d, order, 256, k1, g, eps)
            assert ok, "256 candidates must be enough on this data"
    assert certified > 50


def test_certificate_refuses_when_the_noise_model_is_violated():
    """a candidate whose key is off by more than eps (beyond what g explains) trips the bracket check"""
    T, H = setup()
    q, k1 = 10, 9
    g_all = np.sqrt(((T - H) ** 2).sum(axis=(1, 2)))
    d = np.sqrt(msd_rows(T, q))
    eps = 1e-7
    keys = msd_rows(H, q)
    order = np.argsort(keys, kind="stable")
    keys2 = keys.copy()
    keys2[order[400]] = keys[order[3]]                                  # a far frame reported as near: outside any bracket
    order2 = np.argsort(keys2, kind="stable")
    ok, _ = certificate(keys2, d, order2, 64, k1, g_all[q] + g_all.max(), eps)
    assert not ok
