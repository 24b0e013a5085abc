"""Freeze trajectories of the reference's OneHot and Flat env classes (run in the builder container, where /root/reference
exists):   python tests/golden/make_golden_variants.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_oracle_live_variants import drive, variant_class  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
for kind, size, max_steps, seed, T in (("onehot", 6, 30, 11, 90), ("flat", 8, 100, 13, 150)):
    actions = np.random.RandomState(seed).randint(0, 6, T)
    d = drive(variant_class(kind), dict(size=(size, size), max_steps=max_steps), seed, actions, flat=kind == "flat")
    d.update(actions=actions, size=size, max_steps=max_steps, flat=kind == "flat")
    d["obs"] = d["obs"].astype(np.uint8); d["obs0"] = d["obs0"].astype(np.uint8)
    out = os.path.join(HERE, f"variants_{kind}_{size}x{size}.npz")
    np.savez_compressed(out, **d)
    print(out, os.path.getsize(out), "bytes")
