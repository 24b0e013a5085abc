"""TEST INFRASTRUCTURE ONLY -- build oracle/libcw_oracle.so from oracle/cw_oracle.c with gcc.

There is no ``oracle/_ref``: the reference is pure Python (no C/C++ sources to compile); its
live import (``oracle/ref_shim.py``) only works in the builder container, where it generated tests/golden.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cw_oracle.c")
LIB = os.path.join(HERE, "libcw_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-fPIC", "-shared", "-pthread", "-Wall", "-Wextra", "-o", LIB + ".tmp", SRC]
    subprocess.run(cmd, check=True)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
