"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libcw_oracle.so plus NumPy-array conveniences.

``OracleBatch`` holds a batch of worlds in exactly the SoA layout the CUDA library uses, so a test can upload
the same arrays to the device, run both sides, and compare with ``np.array_equal``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build as _build

_lib = None


class CwoConfig(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("cell_stride", C.c_int32), ("max_steps", C.c_int32),
                ("subset_reward", C.c_int32), ("stacking", C.c_int32), ("n_selected", C.c_int32),
                ("number_of_tasks", C.c_int32), ("selected", C.c_uint8 * 16)]


def cell_stride(H: int, W: int) -> int:
    return (H * W + 15) // 16 * 16


def make_config(H=21, W=21, max_steps=300, subset_reward=False, stacking=True, selected=tuple(range(9)),
                number_of_tasks=None) -> CwoConfig:
    cfg = CwoConfig()
    cfg.H, cfg.W, cfg.cell_stride, cfg.max_steps = H, W, cell_stride(H, W), max_steps
    cfg.subset_reward, cfg.stacking, cfg.n_selected = int(subset_reward), int(stacking), len(selected)
    cfg.number_of_tasks = min(number_of_tasks if number_of_tasks is not None else len(selected), len(selected))
    for i, s in enumerate(selected):
        cfg.selected[i] = s
    return cfg


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.cwo_version.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleBatch:
    """N worlds in the shared SoA layout (see cw_oracle.c header)."""

    def __init__(self, cfg: CwoConfig, N: int, seed: int = 0, env_id_base: int = 0):
        self.cfg, self.N, self.seed, self.env_id_base = cfg, N, seed, env_id_base
        self.lib = load()
        self.grid = np.zeros((N, cfg.cell_stride), np.uint8)
        self.init_grid = np.zeros((N, cfg.cell_stride), np.uint8)
        self.agent = np.zeros(N, np.uint32)
        self.goal = np.zeros(N, np.uint32)
        self.t = np.zeros(N, np.int32)
        self.episode = np.zeros(N, np.uint32)
        self.reward = np.zeros(N, np.int32)
        self.done = np.zeros(N, np.uint8)
        self.stats = np.zeros(24, np.int64)

    # ---- loading / reading compact states -------------------------------------------------------
    def load_state(self, grid, r, c, hold, desired, init_grid=None, achieved=None, t=None):
        H, W = self.cfg.H, self.cfg.W
        g = np.asarray(grid, np.uint8).reshape(self.N, H * W)
        self.grid[:] = 0
        self.grid[:, :H * W] = g
        self.init_grid[:] = self.grid if init_grid is None else 0
        if init_grid is not None:
            self.init_grid[:, :H * W] = np.asarray(init_grid, np.uint8).reshape(self.N, H * W)
        self.agent[:] = (np.asarray(r, np.uint32) | (np.asarray(c, np.uint32) << 8) | (np.asarray(hold, np.uint32) << 16))
        ach = np.zeros(self.N, np.uint32) if achieved is None else np.asarray(achieved, np.uint32)
        self.goal[:] = ach | (np.asarray(desired, np.uint32) << 16)
        self.t[:] = 0 if t is None else t

    @property
    def grid2d(self):
        H, W = self.cfg.H, self.cfg.W
        return self.grid[:, :H * W].reshape(self.N, H, W)

    @property
    def r(self):
        return (self.agent & 0xFF).astype(np.uint8)

    @property
    def c(self):
        return ((self.agent >> 8) & 0xFF).astype(np.uint8)

    @property
    def hold(self):
        return ((self.agent >> 16) & 0xFF).astype(np.uint8)

    @property
    def achieved(self):
        return (self.goal & 0xFFFF).astype(np.uint16)

    @property
    def desired(self):
        return (self.goal >> 16).astype(np.uint16)

    def frame_shape(self):
        return (self.N, 4 * self.cfg.H, 4 * self.cfg.W, 3)

    # ---- entry points -----------------------------------------------------------------------------
    def step(self, actions):
        a = np.ascontiguousarray(actions, np.uint8)
        self.lib.cwo_step(C.byref(self.cfg), _p(self.grid), _p(self.init_grid), _p(self.agent), _p(self.goal), _p(self.t),
                          _p(a), _p(self.reward), _p(self.done), C.c_int64(self.N))
        return self.reward.copy(), self.done.copy()

    def render(self):
        obs = np.empty(self.frame_shape(), np.uint8)
        self.lib.cwo_render(C.byref(self.cfg), _p(self.grid), _p(self.agent), _p(obs), C.c_int64(self.N))
        return obs

    def reset(self, mask=None, with_goal=False):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        goal_obs = np.zeros(self.frame_shape(), np.uint8) if with_goal else None
        self.lib.cwo_reset(C.byref(self.cfg), _p(self.grid), _p(self.init_grid), _p(self.agent), _p(self.goal), _p(self.t),
                           _p(self.episode), _p(m), C.c_uint64(self.seed), C.c_uint64(self.env_id_base), _p(goal_obs),
                           C.c_int64(self.N))
        return goal_obs

    def imagine(self):
        out_grid = np.zeros_like(self.grid)
        out_agent = np.zeros_like(self.agent)
        self.lib.cwo_imagine(C.byref(self.cfg), _p(self.grid), _p(self.agent), _p(self.goal), _p(self.episode),
                             C.c_uint64(self.seed), C.c_uint64(self.env_id_base), _p(out_grid), _p(out_agent), C.c_int64(self.N))
        return out_grid, out_agent

    def step_full(self, actions, auto_reset=True, obs=None, goal_obs=None, stats=True):
        a = np.ascontiguousarray(actions, np.uint8)
        self.lib.cwo_step_full(C.byref(self.cfg), _p(self.grid), _p(self.init_grid), _p(self.agent), _p(self.goal), _p(self.t),
                               _p(self.episode), _p(a), _p(self.reward), _p(self.done), _p(obs), _p(goal_obs),
                               _p(self.stats) if stats else None, C.c_int(int(auto_reset)), C.c_uint64(self.seed),
                               C.c_uint64(self.env_id_base), C.c_int64(self.N))
        return self.reward.copy(), self.done.copy()

    def run_threads(self, actions, render_mode=0, obs=None, nthreads=1):
        a = np.ascontiguousarray(actions, np.uint8)
        K = a.shape[0]
        assert a.shape == (K, self.N)
        self.lib.cwo_run_threads(C.byref(self.cfg), _p(self.grid), _p(self.init_grid), _p(self.agent), _p(self.goal), _p(self.t),
                                 _p(self.episode), _p(a), _p(self.reward), _p(self.done), _p(obs), _p(self.stats),
                                 C.c_int(render_mode), C.c_uint64(self.seed), C.c_uint64(self.env_id_base),
                                 C.c_int64(self.N), C.c_int(K), C.c_int(nthreads))


def set_fixed_pool(grid=None, agent=None):
    """Install (or clear, with None) the fixed_init_state pool used by every subsequent reset.  The arrays must stay
    alive while installed."""
    lib = load()
    if grid is None:
        lib.cwo_set_fixed_pool(None, None, C.c_int64(0))
    else:
        assert grid.dtype == np.uint8 and agent.dtype == np.uint32 and grid.flags.c_contiguous
        lib.cwo_set_fixed_pool(_p(grid), _p(agent), C.c_int64(grid.shape[0]))


def philox(ctr, key):
    lib = load()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib.cwo_philox4x32_10(c, k, out)
    return tuple(out)


def stream_uniform(seed, env_id, episode, n, count):
    lib = load()
    out = np.zeros(count, np.uint32)
    lib.cwo_stream_uniform(C.c_uint64(seed), C.c_uint64(env_id), C.c_uint32(episode), C.c_uint32(n), C.c_int(count), _p(out))
    return out
