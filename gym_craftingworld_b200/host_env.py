"""Host-buffer env: NumPy in, NumPy out, through the ``cw_host_*`` C entry points (include/cw_b200.h).

This is the call a user of the reference makes today -- ``obs, reward, done, info = env.step(action)`` with host
arrays -- for N worlds per call.  The library owns device state, pinned staging and streams; each step copies the
actions host->device, runs the fused step+reset+render launch in slices, and copies reward/done (and the frames)
device->host.  bench.py's ``e2e`` number is measured through this class.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .env import MAX_STEPS, STATE_H, STATE_W, TASK_LIST, make_config


def pinned_empty(shape, dtype):
    """Page-locked NumPy array (backed by a pinned torch tensor kept alive on the array)."""
    t = torch.empty(tuple(shape), dtype=dtype).pin_memory()
    a = t.numpy()
    return a, t


class HostCraftingWorldEnv:
    def __init__(self, num_envs, size=(STATE_W, STATE_H), max_steps=MAX_STEPS, task_list=TASK_LIST,
                 selected_tasks=TASK_LIST, number_of_tasks=None, stacking=True, reward_style=None, *, device=0, seed=0,
                 auto_reset=True, env_id_base=0, return_frames=True, transport="frames"):
        """``transport``: how frames reach host memory when ``return_frames`` -- ``"frames"`` copies every frame over PCIe;
        ``"delta"`` ships a 16-byte delta per world and patches the pinned frame mirror on the host (bit-identical frames)."""
        if transport not in ("frames", "delta"):
            raise ValueError("transport must be 'frames' or 'delta'")
        self.transport = transport
        self.num_envs = int(num_envs)
        self.cfg = make_config(size, max_steps, task_list, selected_tasks, number_of_tasks, stacking, reward_style)
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("gym_craftingworld_b200 needs a CUDA device: there is no CPU path")
        self.device = int(device)
        self.return_frames = bool(return_frames)
        self._h = C.c_void_p()
        flags = (_lib.F_AUTO_RESET if auto_reset else 0) | (_lib.F_DELTA_TRANSPORT if (transport == "delta" and return_frames) else 0)
        _lib.check(self._lib.cw_host_create(C.byref(self.cfg), self.num_envs, self.device, C.c_uint64(int(seed)),
                                            C.c_uint64(int(env_id_base)), flags, C.byref(self._h)), "cw_host_create")
        N, H, W = self.num_envs, self.cfg.H, self.cfg.W
        self.frame_shape = (N, 4 * H, 4 * W, 3)
        self.obs, self._obs_t = pinned_empty(self.frame_shape, torch.uint8) if return_frames else (None, None)
        self.desired_goal, self._goal_t = pinned_empty(self.frame_shape, torch.uint8) if return_frames else (None, None)
        self.reward, self._rew_t = pinned_empty((N,), torch.int32)
        self._done_u8, self._done_t = pinned_empty((N,), torch.uint8)
        self._actions, self._act_t = pinned_empty((N,), torch.uint8)
        # a step is tens of microseconds: resolve the buffer addresses and build the (in-place mutated) outputs once
        self._step_args = (self._h, self._p(self._actions), self._p(self.reward), self._p(self._done_u8), self._p(self.obs))
        self._done = self._done_u8.view(np.bool_)
        self._obs_dict = {"observation": self.obs, "desired_goal": self.desired_goal, "achieved_goal": self.obs}
        self._info = {}

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.num_envs

    @property
    def d2h_bytes_per_step(self) -> int:
        if self.return_frames and self.transport == "delta":
            return self.num_envs * 16           # delta records (+ 72 B per re-seeded world, ~1/300 of the worlds per step)
        return self.num_envs * 5 + (int(np.prod(self.frame_shape)) if self.return_frames else 0)

    def _p(self, a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    def reset(self):
        _lib.check(self._lib.cw_host_reset(self._h, self._p(self.obs), self._p(self.desired_goal)), "cw_host_reset")
        return self._obs_dict

    def step(self, actions):
        """``(obs dict, reward int32[N], done bool[N], info)``; like the reference (``ray.py:194-196, 359-360``) the returned
        arrays are owned by the env and mutated in place by the next call."""
        np.copyto(self._actions, np.asarray(actions).reshape(-1), casting="unsafe")
        rc = self._lib.cw_host_step(*self._step_args)
        if rc:
            _lib.check(rc, "cw_host_step")
        return self._obs_dict, self.reward, self._done, self._info

    def stats(self):
        s = np.zeros(_lib.STATS_LEN, np.int64)
        _lib.check(self._lib.cw_host_stats(self._h, self._p(s)), "cw_host_stats")
        return s

    def close(self):
        if self._h:
            self._lib.cw_host_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
