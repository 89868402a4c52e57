"""Build + load the in-tree CUDA library (``libseqdex_b200.so``).  There is NO fallback: if the
library is missing or the machine has no GPU, creating an env raises."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SEQDEX_B200_LIB") or os.path.join(_HERE, "libseqdex_b200.so")   # override: A/B builds of the same sources
CSRC = os.path.join(_HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false",            # rounding contract of the contact step (csrc/sdx_math.cuh)
              "-Xcompiler", "-fPIC", "-shared"]
_LIB = None


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def build(force=False, verbose=False):
    srcs = sources()
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "seqdex_b200.h")]
    if not force and os.path.exists(SO_PATH) and all(os.path.getmtime(SO_PATH) >= os.path.getmtime(d) for d in deps):
        return SO_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", SO_PATH] + srcs
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return SO_PATH


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(seqdex_b200 has no CPU or PyTorch fallback)")
    L = ctypes.CDLL(SO_PATH)
    L.sdx_last_error.restype = ctypes.c_char_p
    L.sdx_launch_count.restype = ctypes.c_int64
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError("seqdex_b200: " + load().sdx_last_error().decode())
