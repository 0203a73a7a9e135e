"""bb_xtc_to_phipsi / angles_to_sincos (the producers of knn_data's input).
CPU: the oracle's torsion() against known answers and an independent dihedral formula, CLI contract.
GPU: featurize.cu against the oracle on the trp-cage backbone trajectory (float torsions to 1 ulp of the
float result, sin/cos to 2 ulp in double), the tools end to end, and the whole phi-psi -> sin/cos -> knn_data
chain of examples/cluster_phipsi.bash."""
import os
import subprocess

import numpy as np
import pytest

from conftest import DATA, ROOT

BIN = os.path.join(ROOT, "mdsctk_b200", "bin")


@pytest.fixture(scope="module", autouse=True)
def built():
    from mdsctk_b200 import build
    build.build_all()


def dihedral(p):
    b0, b1, b2 = p[0] - p[1], p[2] - p[1], p[3] - p[2]
    b1n = b1 / np.linalg.norm(b1)
    v, w = b0 - np.dot(b0, b1n) * b1n, b2 - np.dot(b2, b1n) * b1n
    return np.arctan2(np.dot(np.cross(b1n, v), w), np.dot(v, w))


def test_oracle_torsion_known_answers_and_independent_formula():
    from oracle import binding as ob
    cis = np.array([[[1, 1, 0], [1, 0, 0], [0, 0, 0], [0, 1, 0], [0, 1, 1], [5, 5, 5]]], dtype=np.float32)
    a = ob.phipsi(cis)      # 6 atoms -> 2 angles: atoms 0-3 and 2-5
    assert a.shape == (1, 2) and abs(a[0, 0]) < 1e-7                        # planar cis = 0
    trans = np.array([[[1, 1, 0], [1, 0, 0], [0, 0, 0], [0, -1, 0], [3, 1, 2], [0, 2, 5]]], dtype=np.float32)
    assert abs(abs(ob.phipsi(trans)[0, 0]) - np.pi) < 1e-6                  # planar trans = +-pi
    plus90 = np.array([[[1, 1, 0], [1, 0, 0], [0, 0, 0], [0, 0, 1], [3, 1, 2], [0, 2, 5]]], dtype=np.float32)
    assert abs(abs(ob.phipsi(plus90)[0, 0]) - np.pi / 2) < 1e-6
    rng = np.random.default_rng(0)
    x = rng.normal(size=(50, 30, 3)).astype(np.float32)
    a = ob.phipsi(x)
    assert a.shape == (50, 18)
    xd = x.astype(np.float64)
    for f in (0, 17, 49):
        for t in range(18):
            s = 3 * (t // 2) + (2 if t % 2 else 0)
            assert abs(a[f, t] - dihedral(xd[f, s:s + 4])) < 5e-5, (f, t)
    sc = ob.sincos(a)
    assert np.array_equal(sc[0::2], np.sin(a).ravel()) and np.array_equal(sc[1::2], np.cos(a).ravel())


def test_featuriser_cli_contract(tmp_path):
    def run(tool, *args):
        p = subprocess.run([os.path.join(BIN, tool), *args], capture_output=True, text=True, cwd=tmp_path)
        return p.returncode, p.stdout
    rc, out = run("bb_xtc_to_phipsi", "-h")
    assert rc == 1 and "usage: bb_xtc_to_phipsi [options]" in out and "--xtc-file" in out and "(=traj.xtc)" in out \
        and "(=phipsi.dat)" in out                                            # bb_xtc_to_phipsi.cpp:60-64
    rc, out = run("angles_to_sincos", "--help")
    assert rc == 1 and "--input-file" in out and "(=phipsi.dat)" in out and "(=sincos.dat)" in out   # angles_to_sincos.cpp:58-62
    rc, out = run("bb_xtc_to_phipsi", "-x", "missing.xtc")
    assert rc == 3 and "xtc-file =    missing.xtc" in out and "output-file = phipsi.dat" in out
    rc, out = run("angles_to_sincos", "-i", "missing.dat")
    assert rc == 3 and "input-file  = missing.dat" in out


@pytest.mark.gpu
def test_gpu_featuriser_matches_oracle(trpcage):
    import mdsctk_b200
    from oracle import binding as ob
    xyz, _mass = trpcage                                   # 1000 frames x 60 backbone atoms (N-CA-C of 20 residues)
    want = ob.phipsi(xyz)
    with mdsctk_b200.KnnContext(0) as ctx:
        ang, sc = ctx.phipsi(xyz)
        assert ang.shape == (1000, 38) and sc.shape == (1000, 76)
        ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
        assert (np.abs(ang - want) <= ulp).all()           # float result within 1 ulp (acos differs between libms)
        assert (ang == want).mean() > 0.9
        assert np.abs(sc.reshape(-1, 2)[:, 0] - np.sin(ang).ravel()).max() < 4.5e-16
        assert np.abs(sc.reshape(-1, 2)[:, 1] - np.cos(ang).ravel()).max() < 4.5e-16
        only = ctx.phipsi(xyz, want_sincos=False)[0]
        assert np.array_equal(only, ang)
        sc2 = ctx.sincos(ang)
        assert np.array_equal(sc2, sc.ravel())
        # ragged: 7 atoms -> natoms/3 = 2 residues -> 2 angles from the first 6 atoms
        a7 = ctx.phipsi(xyz[:5, :7])[0]
        assert np.array_equal(a7, ctx.phipsi(xyz[:5, :6])[0]) and a7.shape == (5, 2)


@pytest.mark.gpu
def test_phipsi_workflow_end_to_end(tmp_path):
    """examples/cluster_phipsi.bash: bb_xtc_to_phipsi -> angles_to_sincos -> knn_data -> make_sysparse."""
    from oracle import binding as ob

    def run(tool, *args):
        p = subprocess.run([os.path.join(BIN, tool), *args], capture_output=True, text=True, cwd=tmp_path)
        assert p.returncode == 0, p.stdout
        return p.stdout
    out = run("bb_xtc_to_phipsi", "-x", os.path.join(DATA, "trp-cage.xtc"), "-s", "fused.dat")
    assert "Wrote 1000 vectors of length 38 (38000 total values)." in out                   # bb_xtc_to_phipsi.cpp:126-128
    out = run("angles_to_sincos")
    assert "Wrote 38000 sin-cos pairs (76000 total values)." in out                         # angles_to_sincos.cpp:124-125
    ang = np.fromfile(tmp_path / "phipsi.dat", dtype=np.float64).reshape(1000, 38)
    sc = np.fromfile(tmp_path / "sincos.dat", dtype=np.float64).reshape(1000, 76)
    assert np.array_equal(sc, np.fromfile(tmp_path / "fused.dat", dtype=np.float64).reshape(1000, 76))
    xyz = ob.read_xtc(os.path.join(DATA, "trp-cage.xtc"))
    assert np.abs(ang - ob.phipsi(xyz)).max() < 3e-7
    run("knn_data", "-k", "12", "-v", "76", "-r", "sincos.dat")
    d, i = ob.knn_data(sc, 12)
    assert np.array_equal(np.fromfile(tmp_path / "indices.dat", dtype=np.int32).reshape(1000, 12), i)
    assert np.array_equal(np.fromfile(tmp_path / "distances.dat", dtype=np.float64).reshape(1000, 12), d)
    run("make_sysparse", "-k", "12")
    raw = (tmp_path / "distances.ssm").read_bytes()
    pcol, irow, val = ob.make_sysparse(i, d)
    assert raw == np.int32(1000).tobytes() + pcol.tobytes() + irow.tobytes() + val.tobytes()
