"""CPU: the NumPy restatement of the AltObs renderer (oracle/compact.py render_alt) against frames frozen from the
reference's CraftingWorldEnvAltObs.render (tests/golden/altobs_8x8.npz), and live where the reference exists."""
import numpy as np
import pytest

from oracle import compact, ref_shim
from tests import golden_util as gu


def test_render_alt_matches_reference_frames():
    src, frame_t, frames = gu.load_altobs()
    d = gu.load(src)
    assert frames.max() > 255                                  # the >255 quirk is in the pin
    for b in range(frames.shape[0]):
        for i, t in enumerate(frame_t):
            img = compact.render_alt(d["grid"][b, t], int(d["r"][b, t]), int(d["c"][b, t]), int(d["hold"][b, t]))
            assert np.array_equal(img, frames[b, i]), (b, t)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference source tree not present")
def test_render_alt_live():
    alt = ref_shim.load_reference_altobs()
    rng = np.random.RandomState(4)
    for size in (4, 9, 21):
        env = alt.CraftingWorldEnvAltObs(size=(size, size))
        for _ in range(20):
            grid = np.where(rng.random_sample((size, size)) < 0.5, rng.randint(1, 9, (size, size)), 0).astype(np.uint8)
            r, c, h = int(rng.randint(size)), int(rng.randint(size)), int(rng.randint(4))
            want = env.render(ref_shim.compact_to_onehot(grid, r, c, h))
            assert np.array_equal(compact.render_alt(grid, r, c, h), want)
