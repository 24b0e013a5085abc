"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference env under a stub ``gym``/``matplotlib``.

``gym`` and ``matplotlib`` are not installed in this image and there is no network, but every piece of
hot-path arithmetic lives in the reference's own files + NumPy; the third-party surface it touches is
metadata only (``gym.spaces``), a base class (``gym.GoalEnv``), and the RNG factory
(``gym.utils.seeding.np_random``; gym <= 0.21 returns a ``np.random.RandomState``).  Attribute uses:
``craftingworld_ray.py:1-3, 8-11, 53, 85-110, 112, 133, 146`` and ``gym_craftingworld/__init__.py:3``.

The reference root is resolved from ``$CW_REFERENCE``, then ``/root/reference`` (the builder container), then
``oracle/_ref`` -- the unmodified package as installed by ``oracle/build_ref.py`` (git-ignored, travels to the GPU box
with the built ``.so`` files).  Callers must still use :func:`reference_available` and fall back to the frozen traces
in ``tests/golden/`` when none of them exists.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

_CANDIDATES = [os.environ.get("CW_REFERENCE", ""), "/root/reference",
               os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]


def reference_root():
    for cand in _CANDIDATES:
        if cand and os.path.isfile(os.path.join(cand, "gym_craftingworld", "envs", "craftingworld_ray.py")):
            return cand
    return None


def reference_available() -> bool:
    return reference_root() is not None


class _Box:
    def __init__(self, low=0, high=1, shape=None, dtype=int):
        self.shape = tuple(shape)
        self.dtype = dtype
        self.low = np.full(self.shape, low, dtype=dtype)
        self.high = np.full(self.shape, high, dtype=dtype)


class _Dict:
    def __init__(self, spaces=None, **kw):
        self.spaces = dict(spaces or {}, **kw)


class _Discrete:
    def __init__(self, n):
        self.n = int(n)
        self._rng = np.random.RandomState()

    def sample(self):
        return int(self._rng.randint(self.n))


def _np_random(seed=None):
    return np.random.RandomState(seed), seed


def install_shim() -> None:
    """Insert stub modules into ``sys.modules`` (idempotent)."""
    if "gym" in sys.modules and getattr(sys.modules["gym"], "_cw_shim", False):
        return

    gym = types.ModuleType("gym")
    gym._cw_shim = True

    class Env:
        metadata = {}

    class GoalEnv(Env):
        pass

    gym.Env, gym.GoalEnv = Env, GoalEnv

    spaces = types.ModuleType("gym.spaces")
    spaces.Box, spaces.Dict, spaces.Discrete = _Box, _Dict, _Discrete
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = _np_random
    utils.seeding = seeding
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.registry = {}

    def register(id, entry_point=None, kwargs=None, **_):
        registration.registry[id] = (entry_point, kwargs or {})

    registration.register = register
    envs.registration = registration
    gym.spaces, gym.utils, gym.envs = spaces, utils, envs

    mpl = types.ModuleType("matplotlib")
    mods = {"gym": gym, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding,
            "gym.envs": envs, "gym.envs.registration": registration, "matplotlib": mpl}
    for sub in ("pyplot", "animation", "patches"):
        m = types.ModuleType("matplotlib." + sub)
        setattr(mpl, sub, m)
        mods["matplotlib." + sub] = m
    for name, mod in mods.items():
        sys.modules.setdefault(name, mod)


def load_reference():
    """Return the reference's ``craftingworld_ray`` module (unmodified source, imported in place)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not present (set $CW_REFERENCE); use the frozen tests/golden traces")
    install_shim()
    if root not in sys.path:
        sys.path.insert(0, root)
    import gym_craftingworld.envs.craftingworld_ray as ray  # noqa: E402
    return ray


def load_reference_altobs():
    """The reference's ``craftingworld_altobs`` module (unregistered AltObs variant)."""
    load_reference()
    import gym_craftingworld.envs.craftingworld_altobs as alt  # noqa: E402
    return alt


# ----------------------------------------------------------------------------------------------------
# converters: reference one-hot int[H,W,12]  <->  compact state
# ----------------------------------------------------------------------------------------------------

def onehot_to_compact(state):
    """one-hot ``int[H,W,12]`` -> ``(grid uint8[H,W], r, c, hold)``; channel layout ``ray.py:605, 784-792``."""
    state = np.asarray(state)
    grid = np.zeros(state.shape[:2], np.uint8)
    for k in range(8):
        grid[state[:, :, k] == 1] = k + 1
    rr, cc = np.where(state[:, :, 8] == 1)
    r, c = int(rr[0]), int(cc[0])
    h = state[r, c, 9:12]
    hold = int(np.argmax(h)) + 1 if h.any() else 0
    return grid, r, c, hold


def compact_to_onehot(grid, r, c, hold):
    grid = np.asarray(grid)
    st = np.zeros(grid.shape + (12,), dtype=int)
    for k in range(8):
        st[:, :, k] = grid == k + 1
    st[r, c, 8] = 1
    if hold:
        st[r, c, 8 + hold] = 1
    return st


def bits_to_mask(vec) -> int:
    return int(sum(int(b) << i for i, b in enumerate(np.asarray(vec).reshape(-1))))


def mask_to_bits(mask: int, n: int = 9):
    return np.array([(mask >> i) & 1 for i in range(n)], dtype=int)


def make_injected_env(ray, grid, r, c, hold, desired_mask, max_steps=300, reward_style=None):
    """Build a reference env and inject an arbitrary compact state into it (SURVEY.md Appendix B.2;
    fields per ``ray.py:176-203``)."""
    H, W = grid.shape
    env = ray.CraftingWorldEnvRay(size=(W, H), max_steps=max_steps, reward_style=reward_style)
    st = compact_to_onehot(grid, r, c, hold)
    env.obs_one_hot = st
    env.agent_pos = ray.Coord(r, c, env.STATE_W - 1, env.STATE_H - 1)
    env.INIT_OBS_VECTOR = st.copy()
    env.desired_goal_vector = mask_to_bits(desired_mask).reshape(1, 9)
    env.achieved_goal_vector = np.zeros((1, 9), dtype=int)
    env.obs_image = env.render(st)
    env.INIT_OBS = env.obs_image.copy()
    env.desired_goal = env.obs_image.copy()
    env.observation = {"observation": env.obs_image, "desired_goal": env.desired_goal,
                       "achieved_goal": env.obs_image, "init_observation": env.INIT_OBS}
    env.step_num = 0
    return env


def read_back(env):
    """(grid, r, c, hold, achieved_mask, pixels uint8) from a live reference env (Appendix B.3)."""
    grid, r, c, hold = onehot_to_compact(env.obs_one_hot)
    assert (r, c) == env.agent_pos.tuple()
    return grid, r, c, hold, bits_to_mask(env.achieved_goal_vector[0]), env.obs_image.astype(np.uint8)
