"""Helpers to read the frozen reference traces in tests/golden (format: tests/golden/make_golden.py)."""
import glob
import os
import zlib

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    """Step/render traces (the AltObs frame file and the OneHot / Flat variant runs have their own layouts and loaders)."""
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(("altobs", "variants_")))


def load_altobs(name="altobs_8x8.npz"):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    return str(z["src"]), z["frame_t"], z["frames"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    d = {k: z[k] for k in z.files}
    for k in ("H", "W", "max_steps", "subset"):
        d[k] = int(d[k])
    return d


def crc(img) -> int:
    return zlib.crc32(np.ascontiguousarray(img, dtype=np.uint8).tobytes()) & 0xFFFFFFFF
