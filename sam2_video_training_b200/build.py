"""Build libsam2b200.so (sm_100a only) in-tree with nvcc.  ``python -m sam2_video_training_b200.build``"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsam2b200.so")
SOURCES = ["abi.cu", "mask_loss.cu", "merge_loss.cu", "attn.cu", "glue.cu", "bank.cu", "mlp.cu", "proj.cu", "lnproj.cu", "wgrad.cu", "memenc.cu", "gemm.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-DNDEBUG", *os.environ.get("SAM2B200_EXTRA_NVCC_FLAGS", "").split()]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsam2b200.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h")) and os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-cudart", "static", "-Wno-deprecated-gpu-targets", "-o", LIB, *objs]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


SELFTEST = os.path.join(HERE, "sam2b200_selftest")


def build_selftest(verbose: bool = True) -> str:
    """Stand-alone GPU self-test binary (no torch) linked against libsam2b200.so."""
    build(verbose=verbose)
    src = os.path.join(CSRC, "selftest.cu")
    if os.path.exists(SELFTEST) and os.path.getmtime(SELFTEST) > max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return SELFTEST
    cmd = [_nvcc(), *NVCC_FLAGS, "-Wno-deprecated-gpu-targets", src, "-o", SELFTEST, "-L", HERE, "-lsam2b200",
           "-Xlinker", "-rpath=$ORIGIN"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return SELFTEST


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    if "--selftest" in sys.argv:
        build_selftest()
    print(LIB)
