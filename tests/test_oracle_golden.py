"""CPU: the NumPy restatement (oracle/compact.py) reproduces every frozen reference trace bit-exactly:
grid, position, held item, achieved vector, reward, done at every step, every frame by CRC and the stored
full frames byte for byte."""
import numpy as np
import pytest

from oracle import compact
from tests import golden_util as gu


@pytest.mark.parametrize("name", gu.golden_files())
def test_compact_matches_reference_trace(name):
    d = gu.load(name)
    cfg = compact.Config(H=d["H"], W=d["W"], max_steps=d["max_steps"], subset_reward=bool(d["subset"]))
    B, T = d["actions"].shape
    fidx = {int(t): i for i, t in enumerate(d["frame_t"])}
    if d["H"] >= 21:                       # pure-Python loop: keep the CPU suite in seconds
        B = min(B, 12)
    for b in range(B):
        s = compact.EnvState(d["grid0"][b].copy(), d["grid0"][b].copy(), int(d["r0"][b]), int(d["c0"][b]),
                             int(d["hold0"][b]), 0, int(d["desired"][b]))
        assert np.array_equal(compact.render(s.grid, s.r, s.c, s.hold), d["frame0"][b])
        for t in range(T):
            reward, done, _ = compact.step_env(s, int(d["actions"][b, t]), cfg)
            where = f"{name} world {b} step {t}"
            assert np.array_equal(s.grid, d["grid"][b, t]), where
            assert (s.r, s.c, s.hold) == (d["r"][b, t], d["c"][b, t], d["hold"][b, t]), where
            assert s.achieved == d["achieved"][b, t], where
            assert reward == d["reward"][b, t] and done == bool(d["done"][b, t]), where
            img = compact.render(s.grid, s.r, s.c, s.hold)
            assert gu.crc(img) == d["frame_crc"][b, t], where
            if t in fidx:
                assert np.array_equal(img, d["frames"][b, fidx[t]]), where


def test_golden_covers_the_quirks():
    """The frozen traces actually exercise the cases Appendix C lists (guards against a vacuous pin)."""
    d = gu.load("dense_5x5_subset.npz")
    ach, des = d["achieved"].astype(int), d["desired"].astype(int)[:, None]
    assert (d["reward"] > 0).any() and (d["reward"] < 0).any()
    assert ((d["done"] == 1) & (d["reward"] < 0)).any()                 # timeout done, and stepping past done
    d = gu.load("dense_8x8.npz")
    ach, des = d["achieved"].astype(int), d["desired"].astype(int)[:, None]
    assert d["hold"].max() == 3 and (d["hold"] == 1).any() and (d["hold"] == 2).any()
    for bit in range(9):
        assert ((ach >> bit) & 1).any(), f"skill bit {bit} never achieved in dense_8x8"
    q = gu.load("quirks_5x5.npz")
    # quirk 1: achieved == desired yet reward -1 (failed move re-evaluates tasks but skips reward)
    assert ((q["achieved"].astype(int) == q["desired"].astype(int)[:, None]) & (q["reward"] < 0)).any()
