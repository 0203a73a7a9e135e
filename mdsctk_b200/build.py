"""Builds libmdsctk_knn.so (the C-ABI CUDA library) and the knn_rms / knn_data tools in-tree.

    python -m mdsctk_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but travels to the
GPU box with the repo snapshot.
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB = os.path.join(PKG, "libmdsctk_knn.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
    "-ccbin", "/usr/bin/g++",
] + os.environ.get("MDSCTK_NVCC_FLAGS", "").split()


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if not force and not _newer(LIB, deps):
        return LIB
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        hdrs = [d for d in deps if not d.endswith(".cu")]
        if force or _newer(o, [s] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-cudart", "static"]
    subprocess.check_call(cmd)
    return LIB


def build_tools(force=False):
    srcs = sorted(glob.glob(os.path.join(HOST, "*.cpp")))
    if not srcs:
        return []
    mk = os.path.join(HOST, "Makefile")
    if os.path.exists(mk):
        subprocess.check_call(["make", "-s", "-C", HOST] + (["-B"] if force else []))
    return [os.path.join(PKG, "bin", t) for t in ("knn_rms", "knn_data", "flatten_xtc", "make_sysparse", "make_gesparse", "bb_xtc_to_phipsi", "angles_to_sincos", "knn_data_sparse", "auto_decomp_sparse", "decomp_sparse")]


def build_all(force=False, verbose=False):
    lib = build_library(force=force, verbose=verbose)
    build_tools(force=force)
    return lib


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
