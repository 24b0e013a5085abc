"""Host-buffer env: NumPy in, NumPy out, through the ``cw_host_*`` C entry points (include/cw_b200.h).

This is the call a user of the reference makes today -- ``obs, reward, done, info = env.step(action)`` with host
arrays -- for N worlds per call.  The library owns device state, pinned staging and streams.  Where the frames go is the
caller's choice (``return_frames`` / ``transport``):

* ``return_frames=False`` -- device consumer: the frames are produced in HBM (four rotating buffers); only the actions (in)
  and reward / done (out) cross PCIe, through mapped pinned memory, and the call returns as soon as reward / done have landed
  (the frames are written behind it, in stream order for a device consumer).  Up to 16 384 worlds a step is a two-launch
  pipeline -- a thread-per-world step launch and the render launch of the snapshot it publishes, on two streams -- so the
  call costs what the frames cost to write (13.7-14.0 us at 4096 worlds) and never waits for the previous step's frames.
* ``transport="delta"`` -- frames current in HOST memory after every call: 16-byte records + host-side patching.
* ``transport="frames"`` -- every rendered frame copied over PCIe.

bench.py's ``e2e`` numbers are measured through this class.
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
        # a step is tens of microseconds: resolve the buffer addresses and build the (in-place mutated) outputs once; the
        # action array is declared to the library, which then reads it in place from the device (it is page-locked)
        self._step_args = (self._h, self._p(self._actions), self._p(self.reward), self._p(self._done_u8), self._p(self.obs))
        _lib.check(self._lib.cw_host_bind_actions(self._h, self._p(self._actions)), "cw_host_bind_actions")
        self._step_fn = self._lib.cw_host_step
        self._many = None
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
        if actions is not self._actions:                       # (callers may also fill ``env.actions`` in place and pass it)
            a = actions if type(actions) is np.ndarray and actions.ndim == 1 else np.asarray(actions).reshape(-1)
            if a.dtype != np.uint8:                            # 260 or -252 must stay the documented no-op, not wrap onto a real action
                a = np.where((a < 0) | (a > 5), 6, a)
            np.copyto(self._actions, a, casting="unsafe")
        rc = self._step_fn(*self._step_args)
        if rc:
            _lib.check(rc, "cw_host_step")
        return self._obs_dict, self.reward, self._done, self._info

    @property
    def actions(self):
        """The env's own page-locked ``uint8[N]`` action array: fill it in place and call ``step(env.actions)`` to skip a copy."""
        return self._actions

    def step_many(self, actions):
        """``K`` consecutive steps on an open-loop action tape ``[K, N]`` in ONE library call (``cw_host_step_many``):
        returns ``(obs dict after the last step, reward int32[K, N], done bool[K, N], info)``.  With ``return_frames=False``
        the K launches are enqueued back to back and the host waits once."""
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        if a.ndim != 2 or a.shape[1] != self.num_envs:
            raise ValueError(f"actions must have shape (K, {self.num_envs})")
        K = a.shape[0]
        if self._many is None or self._many[0][0].shape[0] != K:
            self._many = (pinned_empty((K, self.num_envs), torch.uint8), pinned_empty((K, self.num_envs), torch.int32),
                          pinned_empty((K, self.num_envs), torch.uint8))
        (pa, _), (pr, _), (pd, _) = self._many
        np.copyto(pa, a)
        _lib.check(self._lib.cw_host_step_many(self._h, self._p(pa), K, self._p(pr), self._p(pd), self._p(self.obs)), "cw_host_step_many")
        return self._obs_dict, pr, pd.view(np.bool_), self._info

    def load_state(self, grid=None, agent=None, goal=None, t=None):
        """Inject compact state (``cw_host_load_state``): ``grid uint8[N, H, W]``, packed ``agent`` / ``goal`` words
        ``uint32[N]``, step counters ``int32[N]``; ``None`` leaves a field as is."""
        N, H, W = self.num_envs, self.cfg.H, self.cfg.W
        g = None
        if grid is not None:
            g = np.zeros((N, self.cfg.cell_stride), np.uint8)
            g[:, :H * W] = np.asarray(grid, np.uint8).reshape(N, H * W)
        ag = None if agent is None else np.ascontiguousarray(agent, dtype=np.uint32).reshape(N)
        gl = None if goal is None else np.ascontiguousarray(goal, dtype=np.uint32).reshape(N)
        tt = None if t is None else np.ascontiguousarray(t, dtype=np.int32).reshape(N)
        _lib.check(self._lib.cw_host_load_state(self._h, self._p(g), self._p(ag), self._p(gl), self._p(tt), self._p(self.obs)),
                   "cw_host_load_state")
        return self._obs_dict

    def sync(self):
        """Wait for everything the handle has enqueued (with ``return_frames=False`` the frames of the last step)."""
        _lib.check(self._lib.cw_host_sync(self._h), "cw_host_sync")

    def fetch_frames(self):
        """``return_frames=False``: copy the current device frames to host memory -> ``(obs uint8[N,4H,4W,3], desired_goal)``."""
        obs, goal = np.empty(self.frame_shape, np.uint8), np.empty(self.frame_shape, np.uint8)
        _lib.check(self._lib.cw_host_fetch_frames(self._h, self._p(obs), self._p(goal)), "cw_host_fetch_frames")
        return obs, goal

    def device_frames(self):
        """Device pointer (int) of the frame buffer holding the current observation (``return_frames=False``)."""
        p = C.c_void_p()
        _lib.check(self._lib.cw_host_device_state(self._h, None, C.byref(p)), "cw_host_device_state")
        return p.value

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
