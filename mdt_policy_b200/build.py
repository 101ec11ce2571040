"""Builds ``libmdtb200.so`` (the C-ABI CUDA library, include/mdtb200.h) in-tree with nvcc for sm_100a.

The library has no torch/python dependency: it is compiled straight from ``csrc/engine.cu`` and
loaded with ctypes (``mdt_policy_b200/_lib.py``).  nvcc cross-compiles without a GPU, so this also
serves as the CPU-side "does it build" check (``__graft_entry__.build``).
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmdtb200.so")
SOURCES = [os.path.join(CSRC, "engine.cu")]
HEADERS = [os.path.join(CSRC, "kernels_simt.cuh"), os.path.join(CSRC, "gemm_tcgen05.cuh"), os.path.join(CSRC, "fused_decoder.cuh"), os.path.join(CSRC, "perceiver.cuh"), os.path.join(CSRC, "perceiver_host.cuh"),
           os.path.join(CSRC, "kernels_train.cuh"), os.path.join(CSRC, "ops_train.cuh"),
           os.path.join(CSRC, "kernels_train2.cuh"), os.path.join(CSRC, "ops_train2.cuh"),
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


HASH_PATH = LIB_PATH + ".srchash"      # sha256 of the sources the library was built from (travels with the .so)
LOCK_PATH = LIB_PATH + ".lock"


def source_hash() -> str:
    h = hashlib.sha256()
    for p in SOURCES + HEADERS:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale() -> bool:
    """True when the library is missing or was built from different sources (content hash, not mtimes: a snapshot copied to
    another machine keeps its hash).  A library without a hash file is trusted as long as it exists."""
    if not os.path.exists(LIB_PATH):
        return True
    if not os.path.exists(HASH_PATH):
        return False
    try:
        with open(HASH_PATH) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Builds in-tree.  Safe under torchrun: one process compiles (file lock, private temp file, atomic rename), the others wait
    and find the finished library."""
    if not force and not is_stale():
        return LIB_PATH
    with open(LOCK_PATH, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():      # another process finished the build while we waited for the lock
                return LIB_PATH
            tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
            cmd = [find_nvcc(), *NVCC_FLAGS, "-o", tmp, *SOURCES]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), file=sys.stderr)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr, file=sys.stderr)
            os.replace(tmp, LIB_PATH)
            with open(HASH_PATH + ".tmp", "w") as f:
                f.write(source_hash())
            os.replace(HASH_PATH + ".tmp", HASH_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
