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
