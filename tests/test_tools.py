"""The knn_rms / knn_data / flatten_xtc command-line tools: CLI contract of the reference tools
(options, exit codes, stdout) on CPU; output files against the golden vectors on the GPU, parsed
the way the downstream consumer make_sysparse does (make_sysparse.cpp:220-277)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import DATA, GOLDEN, ROOT

BIN = os.path.join(ROOT, "mdsctk_b200", "bin")


@pytest.fixture(scope="module", autouse=True)
def built():
    from mdsctk_b200 import build
    build.build_all()


def run(tool, *args, cwd=None):
    p = subprocess.run([os.path.join(BIN, tool), *args], capture_output=True, text=True, cwd=cwd, timeout=600)
    rc = p.returncode if p.returncode >= 0 else p.returncode
    return rc, p.stdout


def test_help_returns_1_and_lists_reference_options():
    rc, out = run("knn_rms", "-h")
    assert rc == 1 and "usage: knn_rms [options]" in out          # knn_rms.cpp:93-97
    for opt in ("--threads", "--knn", "--sort", "--nofit", "--block-size", "--topology-file", "--reference-file",
                "--fit-file", "--distance-file", "--index-file"):     # knn_rms.cpp:74-86
        assert opt in out
    rc, out = run("knn_data", "--help")
    assert rc == 1
    for opt in ("--vector-size", "--correlation", "--reference-file", "--knn"):   # knn_data.cpp:70-82
        assert opt in out


def test_missing_required_options_return_minus_1():
    rc, out = run("knn_rms")
    assert rc == 255 and "ERROR: --knn not supplied." in out         # knn_rms.cpp:98-108
    rc, out = run("knn_data", "-r", os.path.join(DATA, "rings.pts"))
    assert rc == 255 and "ERROR: --knn not supplied." in out and "ERROR: --vector-size not supplied." in out


def test_banner_and_option_echo(tmp_path):
    rc, out = run("knn_rms", "--knn=5", "-p", "nope.pdb")
    assert "MDSCTK" in out and "Running with the following options:" in out
    assert "knn =            5" in out and "topology-file =  nope.pdb" in out
    assert rc == 3 and "ERROR" in out


def test_long_option_prefix_and_bool_values(tmp_path):
    rc, out = run("knn_rms", "--kn", "7", "--sor", "false", "--nofit=yes", "-p", "nope.pdb")
    assert "knn =            7" in out and "sort =           0" in out and "nofit =          1" in out
    rc, out = run("knn_rms", "--bogus", "1")
    assert rc == 2 and "unrecognised option" in out


def test_atom_count_mismatch_exits_4(tmp_path):
    mf = tmp_path / "mass.txt"
    mf.write_text("\n".join(["12.0"] * 59))
    rc, out = run("knn_rms", "-k", "3", "-m", str(mf), "-r", os.path.join(DATA, "trp-cage.xtc"))
    assert rc == 4 and "does not match the number of atoms" in out   # knn_rms.cpp:158-178


def test_flatten_xtc_matches_oracle_decoder(tmp_path):
    from oracle import binding as ob
    out = tmp_path / "t.crd"
    rc, txt = run("flatten_xtc", "-x", os.path.join(DATA, "trp-cage.xtc"), "-o", str(out), "-t", "3")
    assert rc == 0 and "Number of frames: 1000" in txt
    got = np.fromfile(out, dtype=np.float32).reshape(-1, 60, 3)
    assert np.array_equal(got, ob.read_xtc(os.path.join(DATA, "trp-cage.xtc")))


def read_like_make_sysparse(dfile, ifile, maxk):
    """make_sysparse.cpp:220-226: nframes = filesize(indices)/4/maxk; rows of maxk doubles / ints."""
    nframes = os.path.getsize(ifile) // 4 // maxk
    assert os.path.getsize(ifile) == nframes * maxk * 4 and os.path.getsize(dfile) == nframes * maxk * 8
    return (np.fromfile(dfile, dtype=np.float64).reshape(nframes, maxk),
            np.fromfile(ifile, dtype=np.int32).reshape(nframes, maxk))


@pytest.mark.gpu
def test_knn_rms_tool_writes_reference_compatible_files(tmp_path):
    g = np.load(os.path.join(GOLDEN, "trpcage_rms_k10.npz"))
    rc, out = run("knn_rms", "-t", "2", "-k", "10", "-p", os.path.join(DATA, "trp-cage.pdb"),
                  "-r", os.path.join(DATA, "trp-cage.xtc"), cwd=tmp_path)      # examples/cluster_rms.bash:51
    assert rc == 0, out
    assert "Block size: 128" in out
    dist, idx = read_like_make_sysparse(tmp_path / "distances.dat", tmp_path / "indices.dat", 10)
    assert dist.shape == (1000, 10)
    assert np.array_equal(idx, g["idx_f64"]) and np.array_equal(idx, g["idx_ref"])
    assert (np.abs(dist - g["dist_f64"]) <= 1e-9 * g["dist_f64"]).all()
    assert (np.abs(dist - g["dist_ref"]) <= 1e-4 * g["dist_ref"]).all()


@pytest.mark.gpu
def test_knn_rms_tool_crd_input_nofit_and_fit_file(tmp_path):
    crd = tmp_path / "t.crd"
    run("flatten_xtc", "-x", os.path.join(DATA, "trp-cage.xtc"), "-o", str(crd))
    g = np.load(os.path.join(GOLDEN, "trpcage_rms_nofit_k10.npz"))
    rc, out = run("knn_rms", "-k", "10", "-n", "true", "-p", os.path.join(DATA, "trp-cage.pdb"), "-r", str(crd),
                  "-d", str(tmp_path / "d.dat"), "-i", str(tmp_path / "i.dat"))
    assert rc == 0, out
    dist, idx = read_like_make_sysparse(tmp_path / "d.dat", tmp_path / "i.dat", 10)
    assert np.array_equal(idx, g["idx_f64"])
    # -f: out-of-sample rows against the landmarks, sorted position 0 still dropped (knn_rms.cpp:284-285)
    from oracle import binding as ob
    xyz = ob.read_xtc(os.path.join(DATA, "trp-cage.xtc"))
    mass = ob.read_masses(os.path.join(DATA, "trp-cage.pdb"))
    xyz[:100].tofile(tmp_path / "fit.crd")
    xyz[100:].tofile(tmp_path / "ref.crd")
    rc, out = run("knn_rms", "-k", "6", "-p", os.path.join(DATA, "trp-cage.pdb"), "-r", str(tmp_path / "ref.crd"),
                  "-f", str(tmp_path / "fit.crd"), "-d", str(tmp_path / "d2.dat"), "-i", str(tmp_path / "i2.dat"))
    assert rc == 0, out
    dist, idx = read_like_make_sysparse(tmp_path / "d2.dat", tmp_path / "i2.dat", 6)
    d, i = ob.knn_rms(xyz[100:], mass, 6, fit=xyz[:100], mode=1)
    assert np.array_equal(idx, i) and (np.abs(dist - d) <= 1e-9 * d).all()


@pytest.mark.gpu
def test_knn_rms_tool_full_matrix_mode(tmp_path):
    from oracle import binding as ob
    xyz = ob.read_xtc(os.path.join(DATA, "trp-cage.xtc"))[:50]
    mass = ob.read_masses(os.path.join(DATA, "trp-cage.pdb"))
    xyz.tofile(tmp_path / "s.crd")
    rc, out = run("knn_rms", "-s", "false", "-p", os.path.join(DATA, "trp-cage.pdb"), "-r", str(tmp_path / "s.crd"),
                  "-d", str(tmp_path / "d.dat"), "-i", str(tmp_path / "i.dat"))
    assert rc == 0, out
    full = np.fromfile(tmp_path / "d.dat", dtype=np.float64).reshape(50, 50)
    want = ob.rms_rows(xyz, mass, xyz, mode=1)
    assert (np.abs(full - want) <= 1e-9 * want + 1e-6).all()


@pytest.mark.gpu
def test_knn_data_tool_bit_identical(tmp_path):
    g = np.load(os.path.join(GOLDEN, "rings_data_k20.npz"))
    rc, out = run("knn_data", "-t", "2", "-k", "20", "-v", "2", "-r", os.path.join(DATA, "rings.pts"),
                  cwd=tmp_path)                                                  # examples/cluster_data.bash:49
    assert rc == 0, out
    assert "Number of reference coordinates: 400" in out
    dist, idx = read_like_make_sysparse(tmp_path / "distances.dat", tmp_path / "indices.dat", 20)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    g = np.load(os.path.join(GOLDEN, "swissroll_data_oos_k10.npz"))
    rc, out = run("knn_data", "-k", "10", "-v", "3", "-r", os.path.join(DATA, "swissroll.pts"),
                  "-f", os.path.join(DATA, "swissroll-outofsample.pts"), cwd=tmp_path)
    assert rc == 0, out
    dist, idx = read_like_make_sysparse(tmp_path / "distances.dat", tmp_path / "indices.dat", 10)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])


def test_topology_masses_two_letter_elements_and_alpha_carbons(tmp_path):
    """PDB / GRO atom names are upper case: " CA " (alpha carbon, column 14) is carbon, "CA  " (column 13) calcium;
    ions and metals get their own masses (element columns 77-78 when present); nothing falls back silently."""
    import subprocess
    from mdsctk_b200 import build
    build.build_tools()
    tool = os.path.join(ROOT, "mdsctk_b200", "bin", "topology_masses")
    pdb = tmp_path / "t.pdb"
    pdb.write_text(
        "ATOM      1  N   MET A   1      27.340  24.430   2.614  1.00  9.67           N  \n"
        "ATOM      2  CA  MET A   1      26.266  25.413   2.842  1.00 10.38           C  \n"
        "ATOM      3 HD11 LEU A   2      26.266  25.413   2.842  1.00 10.38\n"
        "ATOM      4  SD  MET A   1      26.266  25.413   2.842  1.00 10.38\n"
        "HETATM    5 CA    CA A 101      26.266  25.413   2.842  1.00 10.38          CA  \n"
        "HETATM    6 ZN    ZN A 102      26.266  25.413   2.842  1.00 10.38\n"
        "HETATM    7 FE   HEM A 103      26.266  25.413   2.842  1.00 10.38\n"
        "HETATM    8 CL    CL A 104      26.266  25.413   2.842  1.00 10.38\n"
        "HETATM    9 NA    NA A 105      26.266  25.413   2.842  1.00 10.38          NA  \n"
        "HETATM   10 MG    MG A 106      26.266  25.413   2.842  1.00 10.38\n"
        "END\n")
    out = subprocess.run([tool, str(pdb)], capture_output=True, text=True, check=True).stdout.split()
    got = [float(x) for x in out]
    want = [14.0067, 12.0107, 1.0079, 32.065, 40.08, 65.37, 55.847, 35.453, 22.9897, 24.305]
    assert np.allclose(got, want, atol=1e-3), got
    gro = tmp_path / "t.gro"
    gro.write_text("ions\n    4\n    1MET     CA    1   0.100   0.200   0.300\n    2CA      CA    2   0.100   0.200   0.300\n"
                   "    3ZN      ZN    3   0.100   0.200   0.300\n    4ALA    HB1    4   0.100   0.200   0.300\n   1.0 1.0 1.0\n")
    out = subprocess.run([tool, str(gro)], capture_output=True, text=True, check=True).stdout.split()
    assert np.allclose([float(x) for x in out], [12.0107, 40.08, 65.37, 1.0079], atol=1e-3), out
    # the trp-cage topology of the reference's examples: N / CA / C backbone only
    out = subprocess.run([tool, os.path.join(DATA, "trp-cage.pdb")], capture_output=True, text=True, check=True).stdout.split()
    assert len(out) == 60 and sorted(set(out)) == ["12.01070", "14.00670"]


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except (OSError, subprocess.SubprocessError):
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 4])
def test_tools_with_several_gpus_write_identical_files(tmp_path, gpus):
    """knn_rms / knn_data --gpus N: every GPU packs its own shard, NCCL replicates the packed reference set, fit rows
    are sharded (knn_rms.cpp:268-279: rows are independent) -- the files must be byte-identical to the one-GPU run."""
    if _n_gpus() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    outs = {}
    for g in (1, gpus):
        d = tmp_path / f"g{g}"
        d.mkdir()
        rc, out = run("knn_rms", "-k", "10", "-g", str(g), "-p", os.path.join(DATA, "trp-cage.pdb"),
                      "-r", os.path.join(DATA, "trp-cage.xtc"), cwd=d)
        assert rc == 0, out
        if g > 1:
            assert "replicated with NCCL" in out
        rc, out = run("knn_rms", "-k", "10", "-g", str(g), "-p", os.path.join(DATA, "trp-cage.pdb"),
                      "-r", os.path.join(DATA, "trp-cage.xtc"), "-f", os.path.join(DATA, "trp-cage-outofsample.xtc"),
                      "-d", "oos_d.dat", "-i", "oos_i.dat", cwd=d)
        assert rc == 0, out
        rc, out = run("knn_data", "-k", "10", "-v", "3", "-g", str(g), "-r", os.path.join(DATA, "swissroll.pts"),
                      "-d", "sw_d.dat", "-i", "sw_i.dat", cwd=d)
        assert rc == 0, out
        outs[g] = {f: (d / f).read_bytes() for f in ("distances.dat", "indices.dat", "oos_d.dat", "oos_i.dat", "sw_d.dat", "sw_i.dat")}
    for f, blob in outs[1].items():
        assert blob == outs[gpus][f], f
    assert len(outs[1]["oos_i.dat"]) == 10000 * 10 * 4
