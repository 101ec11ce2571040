"""Builds ``libmdtb200.so`` (the C-ABI CUDA library, include/mdtb200.h) in-tree with nvcc for sm_100a.

The library has no torch/python dependency: it is compiled straight from ``csrc/engine.cu`` and
loaded with ctypes (``mdt_policy_b200/_lib.py``).  nvcc cross-compiles without a GPU, so this also
serves as the CPU-side "does it build" check (``__graft_entry__.build``).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmdtb200.so")
SOURCES = [os.path.join(CSRC, "engine.cu")]
HEADERS = [os.path.join(CSRC, "kernels_simt.cuh"), os.path.join(CSRC, "gemm_tcgen05.cuh"), os.path.join(CSRC, "fused_decoder.cuh"),
           os.path.join(CSRC, "kernels_train.cuh"), os.path.join(CSRC, "ops_train.cuh"),
           os.path.join(os.path.dirname(PKG_DIR), "include", "mdtb200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xcompiler", "-fvisibility=hidden",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found (needed to build libmdtb200.so)")
    return nvcc


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS if os.path.exists(p))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH + ".tmp", *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr, file=sys.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
