"""TEST INFRASTRUCTURE ONLY -- install the UNMODIFIED reference package into ``oracle/_ref/`` so that it travels.

The reference is pure Python (nothing to compile), but ``/root/reference`` exists only in the builder container.  This
recipe runs the one offline install the contract allows,

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target oracle/_ref <copy of /root/reference>

(from a copy under ``/tmp``: the source tree is read-only and setuptools writes ``build/`` + ``*.egg-info`` next to
``setup.py``; ``--no-deps`` because ``gym`` and ``matplotlib`` are not in the wheelhouse -- ``oracle/ref_shim.py``
stubs the metadata-only surface of both).  ``oracle/_ref/`` is git-ignored (reference sources never enter the
history) but NOT gpurun-ignored, so the installed package reaches the GPU box like the built ``.so`` files do.
There ``ref_shim.reference_root()`` falls back to it, which puts the real ``CraftingWorldEnvRay`` into
``bench.py --impl reference``, the ``cpu_baseline`` leg and the live GPU-vs-reference tests.

``python -m oracle.build_ref`` / ``__graft_entry__.build()``.  A no-op (keeping what is there) when the source is absent.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("CW_REFERENCE") or "/root/reference"
MARKER = os.path.join(REF_DIR, "gym_craftingworld", "envs", "craftingworld_ray.py")


def installed() -> bool:
    return os.path.isfile(MARKER)


def build(force: bool = False) -> str | None:
    """Returns the install directory, or None when there is neither a source tree nor a previous install."""
    src_marker = os.path.join(SOURCE, "gym_craftingworld", "envs", "craftingworld_ray.py")
    if not os.path.isfile(src_marker):
        return REF_DIR if installed() else None
    if installed() and not force and os.path.getmtime(MARKER) >= os.path.getmtime(src_marker):
        return REF_DIR
    with tempfile.TemporaryDirectory(prefix="cw_refsrc_") as tmp:
        work = os.path.join(tmp, "src")
        shutil.copytree(SOURCE, work, ignore=shutil.ignore_patterns(".git", "docs"))
        shutil.rmtree(REF_DIR, ignore_errors=True)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", REF_DIR, work]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not installed():
            raise RuntimeError("pip install of the reference into oracle/_ref failed:\n" + res.stdout + res.stderr)
    shutil.rmtree(os.path.join(REF_DIR, "tests"), ignore_errors=True)      # find_packages() also picks up the stale tests/
    return REF_DIR


if __name__ == "__main__":
    print(build(force=True))
