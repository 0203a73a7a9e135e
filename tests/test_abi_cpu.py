"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/mdsctk_knn.h declares; without a GPU it fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
import mdsctk_b200
from mdsctk_b200 import api, build


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return api.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mdsctk_knn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdsctk_knn_\w+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mdsctk_knn.h but not exported"
    assert sorted(api.SYMBOLS) == names
    assert lib.mdsctk_knn_abi_version() == 2


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "mdsctk_knn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "torch" not in text.lower() and "at::" not in text and "std::" not in text


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.KnnError) as e:
        mdsctk_b200.KnnContext(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    h = ctypes.c_void_p()
    assert lib.mdsctk_knn_create(ctypes.byref(h), 0) < 0 and not h.value
    assert lib.mdsctk_knn_last_error(None)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mdsctk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower().replace("test infrastructure", ""), f"{f} mentions the oracle"


def test_built_for_sm100a_with_lineinfo():
    flags = " ".join(build.NVCC_FLAGS)
    assert "arch=compute_100a,code=sm_100a" in flags and "-lineinfo" in flags


def test_rms_sweep_layout_fits_for_every_atom_count(lib):
    """Host arithmetic of the version-2 RMSD sweep's shared-memory / TMEM layout (rms_tc2.cu tc2_layout), every atom count the
    kernel accepts: the dynamic shared memory fits, the control block is what is left after operands, ring depth and
    refine-queue size stay in the ranges the kernel was tested with."""
    out = (ctypes.c_int * 12)()
    supported = wide = 0
    for atoms in range(1, 400):
        for want in (0, 1):
            assert lib.mdsctk_knn_debug_rms_layout(atoms, want, out) == 0
            ok, is_wide, nst, bst, qcap, smem, chunks, tmem_units, nkc, nks, ctl, smem_max = list(out)
            a_pad = (atoms + 15) // 16 * 16
            if a_pad > 304:
                assert not ok                                   # the streaming kernel of round 1 takes over
                continue
            assert ok, atoms
            assert nks == a_pad // 16 and 2 * chunks + tmem_units == 3 * nks and tmem_units <= 10
            assert smem == chunks * 8192 + nst * bst + ctl <= smem_max == 227 * 1024
            assert qcap in (32, 20) and qcap * 5 >= 64          # the warp's queue area doubles as a 64-bin histogram
            if is_wide:
                assert want and bst == 9216 and 3 <= nst <= 8 and nkc == (a_pad + 63) // 64
            else:
                assert bst == 4608 and 4 <= nst <= 8 and nkc == (a_pad + 31) // 32 and qcap == 32
            supported += 1
            wide += is_wide
    assert supported == 2 * 304 and wide >= 250                 # wide stages wherever the planes split evenly and three fit
    assert lib.mdsctk_knn_debug_rms_layout(0, 1, out) < 0
