"""Builds the in-tree CUDA library (sm_100a only).  ``python -m mop_truss_marl_b200.build``"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtfem.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
]
SOURCES = ["tfem_family.cu", "tfem_kernels.cu", "tfem_dense.cu", "tfem_capi.cu", "tactor.cu", "trollout.cu", "tpareto.cu"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "tfem.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "tactor.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "trollout.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "tpareto.h"))
    if force or _stale(LIB_PATH, deps):
        nvcc = os.environ.get("NVCC", "nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
        subprocess.run(cmd, check=True)
    return LIB_PATH


PROF_LIB_PATH = os.path.join(LIB_DIR, "libtfem_prof.so")


def build_prof() -> str:
    """development build with the actor kernel's per-warp cycle counters (-DTACTOR_PROF); used through TFEM_LIB by
    scripts/actor_prof.py, never by the product path"""
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run([os.environ.get("NVCC", "nvcc")] + NVCC_FLAGS + ["-DTACTOR_PROF", "-o", PROF_LIB_PATH] + srcs, check=True)
    return PROF_LIB_PATH


PEAKS_SRC = os.path.join(os.path.dirname(HERE), "scripts", "peaks.cu")
PEAKS_BIN = os.path.join(LIB_DIR, "tfem_peaks")


def build_peaks(force: bool = False) -> str:
    """the pipe-peak microbenchmarks (FP64 FMA, DMMA, warp-level HMMA): a stand-alone binary run on the GPU box"""
    os.makedirs(LIB_DIR, exist_ok=True)
    if force or _stale(PEAKS_BIN, [PEAKS_SRC]):
        nvcc = os.environ.get("NVCC", "nvcc")
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-o", PEAKS_BIN, PEAKS_SRC], check=True)
    return PEAKS_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_peaks(force="--force" in sys.argv))
    if "--prof" in sys.argv:
        print(build_prof())
