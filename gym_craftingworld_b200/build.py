"""Build libcw_b200.so in-tree with nvcc for sm_100a (B200).  ``python -m gym_craftingworld_b200.build``.

The shared object is git-ignored but travels to the GPU box with the gpurun snapshot; it is rebuilt only when a
source is newer.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(PKG, "libcw_b200.so")
SOURCES = ["cw_kernels.cu", "cw_host.cu"]
DEPS = SOURCES + ["cw_device.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC,-Wall", "-Xptxas", "-v", "--use_fast_math"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libcw_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    mt = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.join(INCLUDE, "cw_b200.h")]
    return any(os.path.getmtime(d) > mt for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-o", LIB + ".tmp"] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    with open(os.path.join(PKG, "ptxas_info.txt"), "w") as f:      # registers / spills / smem of every kernel
        f.write(res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
