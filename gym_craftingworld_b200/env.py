"""Batched CraftingWorld env: the host-side mirror of the reference's ``CraftingWorldEnvRay``
(``gym_craftingworld/envs/craftingworld_ray.py``, "ray.py") for N independent worlds resident on one B200.

Same constructor keywords (``ray.py:59-60``), same method names and return structure (``reset`` 156-218,
``step`` 301-378, ``render`` 442-520, ``compute_reward`` 757-767, ``seed`` 145-147), batched over a leading
``num_envs`` axis and returned as CUDA tensors.  All arithmetic happens in the hand-written sm_100a kernels behind
the C ABI of ``include/cw_b200.h``; PyTorch only owns the device buffers and the stream.  There is no CPU path.

Differences from the reference, all additive (DESIGN.md "Boundary"):
  * frames are ``uint8`` (the reference returns int64 arrays whose values are all in 0..255 on this render path);
  * ``auto_reset=True`` re-seeds a finished world inside the same launch: the returned reward/done belong to the
    finished episode, the observation to the new one (the reference has no auto-reset, ``gen_info.rst:75-80``);
  * reset randomness is a Philox4x32-10 counter stream keyed by (seed, GLOBAL env id, episode) instead of the
    reference's MT19937 ``RandomState`` -- same distributions, results independent of how worlds are sharded;
  * an out-of-range action is a no-op that still advances ``step_num`` (the reference raises IndexError,
    ``ray.py:308``); pass ``validate_actions=True`` to get the exception (costs a device sync).
"""
from __future__ import annotations

import ctypes as C
import os
from collections.abc import Mapping

import numpy as np
import torch

from . import _lib
from . import spaces

OBJECTS = ['sticks', 'axe', 'hammer', 'rock', 'tree', 'bread', 'house', 'wheat']          # ray.py:21
PICKUPABLE = ['sticks', 'axe', 'hammer']                                                  # ray.py:20
TASK_LIST = ['MakeBread', 'EatBread', 'BuildHouse', 'ChopTree', 'ChopRock', 'GoToHouse', 'MoveAxe', 'MoveHammer',
             'MoveSticks']                                                                # ray.py:40-41
ACTION_NAMES = ['up', 'right', 'down', 'left', 'pickup', 'drop']                          # ray.py:130-131
STATE_W = STATE_H = 21                                                                    # ray.py:43-44
MAX_STEPS = 300                                                                           # ray.py:46
COLORS_N = [(0, 0, 0), (110, 69, 39), (255, 105, 180), (100, 100, 200), (100, 100, 100), (0, 128, 0),
            (205, 133, 63), (197, 91, 97), (240, 230, 140)]                               # ray.py:28-30
FIXED_POOL_ID_BASE = 1 << 62      # Philox stream ids of the fixed_init_state pool (disjoint from env ids)


def make_config(size=(STATE_W, STATE_H), max_steps=MAX_STEPS, task_list=TASK_LIST, selected_tasks=TASK_LIST,
                number_of_tasks=None, stacking=True, reward_style=None) -> _lib.CwConfig:
    """Validate the reference constructor arguments (``ray.py:59-83``) and pack them for the C ABI."""
    W, H = (int(size[0]), int(size[1]))                                                    # ray.py:75
    if W != H:
        raise ValueError(f"size must be square: non-square sizes are broken upstream (IndexError), got {size}")
    if not (3 <= H <= _lib.MAX_SIDE):
        raise ValueError(f"size must be within 3..{_lib.MAX_SIDE}, got {size}")
    if int(max_steps) < 1:
        raise ValueError("max_steps must be >= 1")
    task_list, selected_tasks = list(task_list), list(selected_tasks)
    if not (9 <= len(task_list) <= 16):
        raise ValueError("task_list must name the 9 skills (eval_task_edit writes bits 0..8, ray.py:657-702)")
    if not (1 <= len(selected_tasks) <= 9):
        raise ValueError("selected_tasks must hold 1..9 tasks")
    for name in selected_tasks:
        if name not in task_list:
            raise ValueError(f"selected task {name!r} is not in task_list")                # ray.py:174 (.index)
    n_tasks = len(selected_tasks) if number_of_tasks is None else int(number_of_tasks)     # ray.py:79
    n_tasks = min(n_tasks, len(selected_tasks))                                            # ray.py:80-81
    if n_tasks < 1:
        raise ValueError("number_of_tasks must be >= 1")
    cfg = _lib.CwConfig()
    cfg.H, cfg.W, cfg.cell_stride, cfg.max_steps = H, W, _lib.cell_stride(H, W), int(max_steps)
    cfg.subset_reward = 0 if reward_style is None else 1                                   # ray.py:71-74
    cfg.stacking = 1 if stacking is True else 0                                            # ray.py:169 ("is True")
    cfg.n_selected, cfg.number_of_tasks = len(selected_tasks), n_tasks
    for i, name in enumerate(selected_tasks):
        cfg.selected[i] = task_list.index(name)
    return cfg


class GoalInfo(Mapping):
    """``info`` of ``step`` (``ray.py:376-378``): 9-vectors unpacked lazily from the packed device word."""
    _KEYS = ("task_success", "desired_goal", "achieved_goal")

    def __init__(self, env):
        self._env = env

    def __getitem__(self, k):
        if k in ("task_success", "achieved_goal"):
            return self._env.achieved_goal_vector
        if k == "desired_goal":
            return self._env.desired_goal_vector
        if k == "achieved_mask":
            return self._env.achieved_mask
        if k == "desired_mask":
            return self._env.desired_mask
        raise KeyError(k)

    def __iter__(self):
        return iter(self._KEYS)

    def __len__(self):
        return len(self._KEYS)


class OneHotObs(Mapping):
    """Observation dict of the one-hot family (``carftingworld_onehot.py:203, 310, 369-371``): values are expanded from
    the compact device state on first access after each step/reset (``achieved_goal is observation``, as upstream)."""
    _KEYS = ("observation", "desired_goal", "achieved_goal", "init_observation")

    def __init__(self, env):
        self._env, self._version, self._cache = env, env._obs_version, {}

    def __getitem__(self, k):
        env = self._env
        if self._version != env._obs_version:
            self._version, self._cache = env._obs_version, {}
        key = "observation" if k == "achieved_goal" else k
        if key not in self._KEYS:
            raise KeyError(k)
        if key not in self._cache:
            if key == "observation":
                self._cache[key] = env.onehot()
            elif key == "desired_goal":
                self._cache[key] = env.onehot(grid=env.goal_grid, agent=env.goal_agent)
            else:
                self._cache[key] = env.onehot(grid=env.init_grid, agent=env.init_agent)
        return self._cache[key]

    def __iter__(self):
        return iter(self._KEYS)

    def __len__(self):
        return len(self._KEYS)


class _NoGuard:
    """``with`` target when the env's device is already current (``torch.cuda.device(...)`` costs 3-4 us per use)."""
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)     # the current stream's handle without building a Stream object


class CompactObs(Mapping):
    """Observation dict of ``obs_mode='compact'``: zero-copy views of the device state; the two goal masks are unpacked from the
    packed goal word ON ACCESS -- ``step`` itself must launch nothing but the step kernel (three tiny elementwise kernels per
    step between two step launches cost more than the step: 5-6 us of a 10 us period at 65 536 worlds, measured)."""
    _KEYS = ("observation", "agent", "desired_goal", "achieved_goal", "init_observation")

    def __init__(self, env):
        self._env = env

    def __getitem__(self, k):
        env = self._env
        if k == "observation":
            return env.grid_view
        if k == "agent":
            return env.agent
        if k == "desired_goal":
            return env.desired_mask
        if k == "achieved_goal":
            return env.achieved_mask
        if k == "init_observation":
            return env.init_grid
        raise KeyError(k)

    def __iter__(self):
        return iter(self._KEYS)

    def __len__(self):
        return len(self._KEYS)


class BatchedCraftingWorldEnv:
    """N independent CraftingWorld worlds on one GPU behind the reference's Env surface."""

    metadata = {'render.modes': ['human', 'Non']}                                          # ray.py:57

    def __init__(self, num_envs, size=(STATE_W, STATE_H), fixed_init_state=0, max_steps=MAX_STEPS, store_gif=False,
                 render_save_rate=1, task_list=TASK_LIST, selected_tasks=TASK_LIST, number_of_tasks=None, stacking=True,
                 reward_style=None, *, device=None, seed=None, auto_reset=True, obs_mode="pixels", env_id_base=0,
                 goal_images=True, obs_buffers=1, validate_actions=False, collect_stats=True, render="full"):
        if store_gif:
            raise NotImplementedError("GIF recording (ray.py:565-597, 769-782) is a host-side debugging side channel; "
                                      "out of scope (DESIGN.md)")
        if obs_mode not in ("pixels", "compact", "onehot"):
            raise ValueError("obs_mode must be 'pixels', 'compact' or 'onehot'")
        if render not in ("full", "incremental"):
            raise ValueError("render must be 'full' (every frame re-expanded every step) or 'incremental' (render_edit)")
        if render == "incremental" and (obs_mode != "pixels" or int(obs_buffers) != 1):
            raise ValueError("render='incremental' patches ONE persistent frame buffer: needs obs_mode='pixels', obs_buffers=1")
        self.render_mode = render
        self.num_envs = int(num_envs)
        if self.num_envs < 1:
            raise ValueError("num_envs must be >= 1")
        self.cfg = make_config(size, max_steps, task_list, selected_tasks, number_of_tasks, stacking, reward_style)
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("gym_craftingworld_b200 needs a CUDA device: there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("gym_craftingworld_b200 needs a CUDA device: there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._dev_index = int(self.device.index)
        self.STATE_W, self.STATE_H = self.cfg.W, self.cfg.H
        self.MAX_STEPS = self.cfg.max_steps
        self.task_list, self.selected_tasks = list(task_list), list(selected_tasks)
        self.number_of_tasks, self.stacking = self.cfg.number_of_tasks, bool(self.cfg.stacking)
        self.fixed_init_state = int(fixed_init_state)
        self.auto_reset, self.obs_mode = bool(auto_reset), obs_mode
        self.goal_images = bool(goal_images) and obs_mode == "pixels"
        self.validate_actions = bool(validate_actions)
        self.collect_stats = bool(collect_stats)
        self.env_id_base = int(env_id_base)
        self.compute_reward = self.compute_reward_equal if reward_style is None else self.compute_reward_subset

        N, H, W, dev = self.num_envs, self.cfg.H, self.cfg.W, self.device
        pw, ph = 4 * W, 4 * H
        img = spaces.Box(0, 255, (pw, ph, 3), np.uint8)
        self.observation_space = spaces.Dict(dict(observation=img, desired_goal=img, achieved_goal=img,
                                                  init_observation=img))                  # ray.py:85-92
        vec = spaces.Box(0, 1, (W, H, len(OBJECTS) + 1 + len(PICKUPABLE)), np.uint8)
        gvec = spaces.Box(0, 1, (1, len(self.task_list)), np.uint8)
        self.observation_vector_space = spaces.Dict(dict(observation=vec, desired_goal=gvec, achieved_goal=gvec,
                                                         init_observation=vec))           # ray.py:94-110
        self.ACTIONS = list(ACTION_NAMES)
        self.action_space = spaces.Discrete(len(self.ACTIONS))                             # ray.py:133

        # ---- device-resident SoA state (layout: include/cw_b200.h) -----------------------------------
        stride = self.cfg.cell_stride
        self.grid = torch.zeros((N, stride), dtype=torch.uint8, device=dev)
        self.init_grid = torch.zeros((N, stride), dtype=torch.uint8, device=dev)
        self.agent = torch.zeros(N, dtype=torch.int32, device=dev)
        self.goal = torch.zeros(N, dtype=torch.int32, device=dev)
        self.t = torch.zeros(N, dtype=torch.int32, device=dev)
        self.episode = torch.zeros(N, dtype=torch.int32, device=dev)
        self.reward = torch.zeros(N, dtype=torch.int32, device=dev)
        self._done_u8 = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.done = self._done_u8.view(torch.bool)
        self.stats_raw = torch.zeros((_lib.STATS_REPLICAS, _lib.STATS_LEN), dtype=torch.int64, device=dev)
        self.frame_shape = (N, ph, pw, 3)
        self._obs_ring, self._ring_pos = [], 0
        self.obs = self.desired_goal = self.init_obs = None
        if obs_mode == "pixels":
            self._obs_ring = [torch.zeros(self.frame_shape, dtype=torch.uint8, device=dev) for _ in range(max(1, int(obs_buffers)))]
            self.obs = self._obs_ring[0]
            if self.goal_images:
                self.desired_goal = torch.zeros(self.frame_shape, dtype=torch.uint8, device=dev)
                self.init_obs = torch.zeros(self.frame_shape, dtype=torch.uint8, device=dev)
        self.goal_grid = self.goal_agent = None
        # agent word at reset = the agent / holding channels of INIT_OBS_VECTOR (ray.py:183); 4 bytes per world, kept in every mode
        self.init_agent = torch.zeros(N, dtype=torch.int32, device=dev)
        if obs_mode == "onehot":                  # compact imagined goal state (one-hot family)
            self.goal_grid = torch.zeros((N, stride), dtype=torch.uint8, device=dev)
            self.goal_agent = torch.zeros(N, dtype=torch.int32, device=dev)
        # compact step path: pre-drawn reset records + their refill queue (cw_prefill_resets; include/cw_b200.h CwState)
        self.reset_rec = self.reset_list = None
        if obs_mode == "compact" and self.auto_reset and not self.fixed_init_state and self.num_envs <= 0x3FFFFFFF:
            self.reset_rec = torch.zeros((N, 8), dtype=torch.int32, device=dev)
            self.reset_list = torch.zeros(4 + 2 * N, dtype=torch.int32, device=dev)
        self._records_fresh = False
        self._obs_version = 0
        self._chain = None                        # chain words of cw_step_render_chained (allocated on first use)
        self._edit_scratch = None                 # work list of cw_step_render_edit (allocated on first use)
        self._fixed_grid = self._fixed_agent = None
        self._seed = None
        self._state = _lib.CwState()
        self._cfg_ref, self._state_ref = C.byref(self.cfg), C.byref(self._state)   # (both structs live as long as the env: built once)
        self.seed(seed)
        self._info = GoalInfo(self)
        self._bits = torch.arange(len(self.task_list), dtype=torch.int32, device=dev)
        self._is_reset = False

    # ------------------------------------------------------------------------------------------------
    def seed(self, seed=None):
        """``seed`` (``ray.py:145-147``): sets the Philox key; returns ``[seed]``."""
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little") >> 1
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._records_fresh = False               # pre-drawn resets belong to the old key
        if self.fixed_init_state:
            self._generate_fixed_states(self.fixed_init_state)
        self._refresh_state_struct()
        return [self._seed]

    def _refresh_state_struct(self):
        s = self._state
        s.grid, s.init_grid = self.grid.data_ptr(), self.init_grid.data_ptr()
        s.agent, s.goal, s.t, s.episode = self.agent.data_ptr(), self.goal.data_ptr(), self.t.data_ptr(), self.episode.data_ptr()
        s.n, s.seed, s.env_id_base = self.num_envs, self._seed, self.env_id_base
        if self._fixed_grid is not None:
            s.fixed_grid, s.fixed_agent, s.n_fixed = self._fixed_grid.data_ptr(), self._fixed_agent.data_ptr(), self.fixed_init_state
        else:
            s.fixed_grid, s.fixed_agent, s.n_fixed = None, None, 0
        s.goal_grid, s.goal_agent, s.init_agent = self._ptr(self.goal_grid), self._ptr(self.goal_agent), self._ptr(self.init_agent)
        s.reset_rec, s.reset_list = self._ptr(getattr(self, "reset_rec", None)), self._ptr(getattr(self, "reset_list", None))

    def _generate_fixed_states(self, n):
        """``generate_fixed_states`` (``ray.py:149-154``): pre-sample ``n`` worlds with ``sample_state``."""
        dev, stride = self.device, self.cfg.cell_stride
        pool = _lib.CwState()
        g = torch.zeros((n, stride), dtype=torch.uint8, device=dev)
        ig = torch.zeros_like(g)
        ag, gl, t, ep = (torch.zeros(n, dtype=torch.int32, device=dev) for _ in range(4))
        pool.grid, pool.init_grid, pool.agent, pool.goal = g.data_ptr(), ig.data_ptr(), ag.data_ptr(), gl.data_ptr()
        pool.t, pool.episode, pool.n, pool.seed, pool.env_id_base = t.data_ptr(), ep.data_ptr(), n, self._seed, FIXED_POOL_ID_BASE
        pool.n_fixed = 0
        pool.goal_grid = pool.goal_agent = pool.init_agent = pool.reset_rec = pool.reset_list = None
        with self._guard():
            _lib.check(self._lib.cw_reset(C.byref(self.cfg), C.byref(pool), None, None, None, None, self._stream()), "cw_reset(pool)")
        self._fixed_grid, self._fixed_agent = g, ag

    def _stream(self):
        if _raw_stream is not None:
            return _raw_stream(self._dev_index)
        return torch.cuda.current_stream(self.device).cuda_stream

    def _guard(self):
        """Device guard for a launch: a no-op when the env's device is the current one (the usual case)."""
        return _NO_GUARD if torch.cuda.current_device() == self._dev_index else torch.cuda.device(self.device)

    # ---- reference-style attributes ------------------------------------------------------------------
    @property
    def stats(self):
        """``int64[24]`` finished-episode statistics of this rank (the device keeps 16 replicas to spread the atomics)."""
        return self.stats_raw.sum(dim=0)

    @property
    def achieved_mask(self):
        return self.goal & 0xFFFF

    @property
    def desired_mask(self):
        return (self.goal >> 16) & 0xFFFF

    @property
    def achieved_goal_vector(self):
        """``uint8[N, len(task_list)]`` (the reference's ``int[1, 9]`` per world, ``ray.py:176``)."""
        return ((self.achieved_mask.unsqueeze(1) >> self._bits) & 1).to(torch.uint8)

    @property
    def desired_goal_vector(self):
        return ((self.desired_mask.unsqueeze(1) >> self._bits) & 1).to(torch.uint8)

    @property
    def step_num(self):
        return self.t

    @property
    def ep_no(self):
        return (self.episode - 1).clamp_(min=0)

    @property
    def agent_pos(self):
        """``int32[N, 2]`` (row, col)."""
        return torch.stack((self.agent & 0xFF, (self.agent >> 8) & 0xFF), dim=1)

    @property
    def holding(self):
        return (self.agent >> 16) & 0xFF

    @property
    def grid_view(self):
        """Zero-copy ``uint8[N, H, W]`` view of the object-code grid (the compact observation)."""
        H, W = self.cfg.H, self.cfg.W
        return self.grid[:, :H * W].view(self.num_envs, H, W)

    def onehot(self, init=False, grid=None, agent=None):
        """One-hot ``uint8[N, H, W, 12]`` state (``obs_one_hot`` / ``INIT_OBS_VECTOR``, ``ray.py:605-613, 183``) of the
        current worlds, or of the given compact ``grid uint8[N, cell_stride]`` / ``agent int32[N]`` tensors."""
        out = torch.empty((self.num_envs, self.cfg.H, self.cfg.W, 12), dtype=torch.uint8, device=self.device)
        src = grid if grid is not None else (self.init_grid if init else self.grid)
        ag = agent if agent is not None else self.agent
        with self._guard():
            _lib.check(self._lib.cw_onehot(C.byref(self.cfg), src.data_ptr(), ag.data_ptr(), out.data_ptr(),
                                           self.num_envs, self._stream()), "cw_onehot")
        return out

    @property
    def observation_vector(self):
        """``ray.py:185-187``: one-hot state, 9-bit goal vectors and ``INIT_OBS_VECTOR`` (object codes of ``init_grid`` plus
        the agent / holding channels as they were at reset, ``init_agent``)."""
        return {"observation": self.onehot(), "desired_goal": self.desired_goal_vector,
                "achieved_goal": self.achieved_goal_vector, "init_observation": self.onehot(grid=self.init_grid, agent=self.init_agent)}

    # ---- observations ----------------------------------------------------------------------------------
    def _observation(self):
        if self.obs_mode == "pixels":
            return {"observation": self.obs, "desired_goal": self.desired_goal, "achieved_goal": self.obs,
                    "init_observation": self.init_obs}                                      # ray.py:194-196
        if self.obs_mode == "onehot":
            return OneHotObs(self)
        return CompactObs(self)

    @property
    def observation(self):
        return self._observation()

    def _ptr(self, t):
        return None if t is None else t.data_ptr()

    def _stats_ptr(self):
        return self.stats_raw.data_ptr() if self.collect_stats else None

    # ---- reset / step ------------------------------------------------------------------------------------
    def reset(self, mask=None):
        """``reset`` (``ray.py:156-218``) for all worlds, or for ``mask`` (bool/uint8 ``[N]``) only."""
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            if m.shape != (self.num_envs,):
                raise ValueError(f"mask must have shape ({self.num_envs},)")
        with self._guard():
            _lib.check(self._lib.cw_reset(C.byref(self.cfg), C.byref(self._state), self._ptr(m), self._ptr(self.obs),
                                          self._ptr(self.desired_goal), self._ptr(self.init_obs), self._stream()), "cw_reset")
        self._prefill_resets()
        self._is_reset = True
        self._obs_version += 1
        return self._observation()

    def _prefill_resets(self):
        """Draw every world's NEXT reset ahead of time (compact step path); a no-op without the record buffers."""
        if self.reset_rec is not None:
            with self._guard():
                _lib.check(self._lib.cw_prefill_resets(C.byref(self.cfg), C.byref(self._state), self._stream()), "cw_prefill_resets")
        self._records_fresh = True

    def _as_actions(self, actions):
        a = actions
        if (type(a) is torch.Tensor and a.dtype is torch.uint8 and a.device == self.device and a.is_contiguous()
                and not self.validate_actions):
            return a                                               # the usual case: a uint8 tensor on the env's device
        if not isinstance(a, torch.Tensor):
            a = torch.as_tensor(np.asarray(a).reshape(-1), device=self.device)
        if a.device != self.device:
            a = a.to(self.device)
        if self.validate_actions and bool(((a < 0) | (a >= len(self.ACTIONS))).any()):
            raise IndexError("action out of range [0, 6)")                                  # ray.py:308
        return self._to_u8_actions(a)

    @staticmethod
    def _to_u8_actions(a):
        """uint8 action codes; an out-of-range value of a wider dtype (260, -252, ...) must stay out of range -- the documented
        no-op -- instead of wrapping modulo 256 onto a real action."""
        if a.dtype != torch.uint8:
            a = torch.where((a < 0) | (a > 5), 6, a).to(torch.uint8)
        return a.contiguous()

    def step(self, actions, chain_pos=None):
        """``step`` (``ray.py:301-378``) for all worlds: ``(obs dict, reward int32[N], done bool[N], info)``.
        One kernel launch; asynchronous on the current stream (CUDA-graph capturable).

        ``chain_pos`` (pixel or compact observations) declares an OPEN-LOOP run of steps -- an action tape that exists before
        the run starts, e.g. the K steps captured into one CUDA graph: pass 0, 1, 2, ... for consecutive calls with
        nothing else enqueued on the stream in between.  Launch ``i > 0`` then follows launch ``i - 1`` by per-group
        dataflow (``cw_step_render_chained``; compact observations: per warp of 32 worlds, ``cw_step_chained``) instead of
        waiting for its whole grid, so its work overlaps the draining frame stores of the previous step.  Results are identical; a closed loop (actions computed from the previous
        observation) must use ``chain_pos=None``."""
        a = self._as_actions(actions)
        if a.shape != (self.num_envs,):
            raise ValueError(f"actions must have shape ({self.num_envs},), got {tuple(a.shape)}")
        if not self._records_fresh and chain_pos is None:      # (never inside a chain: nothing may sit between two positions)
            self._prefill_resets()
        if chain_pos is not None and (self.obs_mode == "onehot" or not 0 <= int(chain_pos) < _lib.CHAIN_MAX_POS):
            raise ValueError(f"chain_pos needs obs_mode 'pixels' or 'compact' and 0 <= chain_pos < {_lib.CHAIN_MAX_POS}")
        flags = _lib.F_AUTO_RESET if self.auto_reset else 0
        with self._guard():
            if self.obs_mode == "pixels":
                if len(self._obs_ring) > 1:
                    self._ring_pos = (self._ring_pos + 1) % len(self._obs_ring)
                    self.obs = self._obs_ring[self._ring_pos]
                if self.render_mode == "incremental":         # the reference's render_edit: patch the <= 2 changed cells
                    if chain_pos is not None:
                        raise ValueError("chain_pos applies to render='full' only")
                    if self._edit_scratch is None:
                        self._edit_scratch = torch.zeros(self.num_envs + 2, dtype=torch.int32, device=self.device)
                    rc = self._lib.cw_step_render_edit(self._cfg_ref, self._state_ref, a.data_ptr(), self.reward.data_ptr(),
                                                       self._done_u8.data_ptr(), self.obs.data_ptr(), self._ptr(self.desired_goal),
                                                       self._ptr(self.init_obs), self._stats_ptr(), flags,
                                                       self._edit_scratch.data_ptr(), self._stream())
                    _lib.check(rc, "cw_step_render_edit")
                    self._obs_version += 1
                    return self._observation(), self.reward, self.done, self._info
                if chain_pos is not None:
                    if self._chain is None:
                        self._chain = torch.zeros(_lib.CHAIN_MAX_POS + self.num_envs, dtype=torch.int32, device=self.device)
                    rc = self._lib.cw_step_render_chained(
                        self._cfg_ref, self._state_ref, a.data_ptr(), self.reward.data_ptr(), self._done_u8.data_ptr(),
                        self.obs.data_ptr(), self._ptr(self.desired_goal), self._ptr(self.init_obs), self._stats_ptr(), flags,
                        self._chain.data_ptr(), int(chain_pos), len(self._obs_ring), self._stream())
                    _lib.check(rc, "cw_step_render_chained")
                    self._obs_version += 1
                    return self._observation(), self.reward, self.done, self._info
                rc = self._lib.cw_step_render(self._cfg_ref, self._state_ref, a.data_ptr(), self.reward.data_ptr(),
                                              self._done_u8.data_ptr(), self.obs.data_ptr(), self._ptr(self.desired_goal),
                                              self._ptr(self.init_obs), self._stats_ptr(), flags, self._stream())
            elif self.obs_mode == "onehot":       # no pixels, but resets must produce the imagined goal state
                rc = self._lib.cw_step_render(self._cfg_ref, self._state_ref, a.data_ptr(), self.reward.data_ptr(),
                                              self._done_u8.data_ptr(), None, None, None, self._stats_ptr(), flags, self._stream())
            elif chain_pos is not None:               # compact observations, open-loop run: launches linked per warp (cw_step_chained)
                if self._chain is None:
                    self._chain = torch.zeros(_lib.CHAIN_MAX_POS + self.num_envs, dtype=torch.int32, device=self.device)
                rc = self._lib.cw_step_chained(self._cfg_ref, self._state_ref, a.data_ptr(), self.reward.data_ptr(),
                                               self._done_u8.data_ptr(), self._stats_ptr(), flags, self._chain.data_ptr(),
                                               int(chain_pos), self._stream())
            else:
                rc = self._lib.cw_step(self._cfg_ref, self._state_ref, a.data_ptr(), self.reward.data_ptr(),
                                       self._done_u8.data_ptr(), self._stats_ptr(), flags, self._stream())
        _lib.check(rc, "cw_step")
        self._obs_version += 1
        return self._observation(), self.reward, self.done, self._info

    def frame_policy(self, obs=None, out=None):
        """A stand-in device consumer (``cw_frame_policy``): reads every byte of every frame of ``obs`` (default: the current
        observation) and returns ``uint8[N]`` actions derived from the pixels -- closes the loop the way a policy network
        does, for measurements and tests.  Asynchronous on the current stream, CUDA-graph capturable."""
        obs = self.obs if obs is None else obs
        if out is None:
            out = torch.empty(self.num_envs, dtype=torch.uint8, device=self.device)
        with self._guard():
            _lib.check(self._lib.cw_frame_policy(C.byref(self.cfg), obs.data_ptr(), self.num_envs, out.data_ptr(), self._stream()),
                       "cw_frame_policy")
        return out

    def rollout(self, actions, return_trace=True):
        """K steps in ONE launch on an open-loop action tape ``uint8[K, N]`` (compact observations only).
        Returns ``(reward int32[K,N], done bool[K,N])`` or ``None`` when ``return_trace`` is False."""
        a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(np.asarray(actions), device=self.device)
        a = self._to_u8_actions(a.to(device=self.device))
        if a.dim() != 2 or a.shape[1] != self.num_envs:
            raise ValueError(f"actions must have shape (K, {self.num_envs})")
        K = a.shape[0]
        if not self._records_fresh:
            self._prefill_resets()
        rew = dn = None
        if return_trace:
            rew = torch.empty((K, self.num_envs), dtype=torch.int32, device=self.device)
            dn = torch.empty((K, self.num_envs), dtype=torch.uint8, device=self.device)
        flags = _lib.F_AUTO_RESET if self.auto_reset else 0
        with self._guard():
            _lib.check(self._lib.cw_rollout(C.byref(self.cfg), C.byref(self._state), a.data_ptr(), self._ptr(rew), self._ptr(dn),
                                            self._stats_ptr(), K, flags, self._stream()), "cw_rollout")
        return (rew, dn.view(torch.bool)) if return_trace else None

    def render(self, state=None, mode="Non", tile_size=4):
        """``render`` (``ray.py:442-520``).  ``state=None`` renders the current worlds; otherwise ``state`` is
        ``(grid uint8[M,H,W], r[M], c[M], hold[M])`` and a fresh ``uint8[M,4H,4W,3]`` tensor is returned."""
        if tile_size != 4:
            raise ValueError("tile_size is fixed at 4 (the reference ignores the argument, ray.py:478-479)")
        H, W = self.cfg.H, self.cfg.W
        if state is None:
            grid, agent, M = self.grid, self.agent, self.num_envs
            out = self.obs if self.obs is not None else torch.empty(self.frame_shape, dtype=torch.uint8, device=self.device)
        else:
            g, r, c, h = state
            g = torch.as_tensor(g, device=self.device).to(torch.uint8).reshape(-1, H * W)
            M = g.shape[0]
            grid = torch.zeros((M, self.cfg.cell_stride), dtype=torch.uint8, device=self.device)
            grid[:, :H * W] = g
            r, c, h = (torch.as_tensor(x, device=self.device).to(torch.int32).reshape(-1) for x in (r, c, h))
            agent = (r | (c << 8) | (h << 16)).contiguous()
            out = torch.empty((M, 4 * H, 4 * W, 3), dtype=torch.uint8, device=self.device)
        with self._guard():
            _lib.check(self._lib.cw_render(C.byref(self.cfg), grid.data_ptr(), agent.data_ptr(), out.data_ptr(), M, self._stream()),
                       "cw_render")
        return out

    # ---- reward (host-visible restatement for callers that relabel goals, e.g. HER) ------------------------
    def _to_mask(self, g):
        """Accept packed masks ``int[N]`` or bit vectors ``[..., len(task_list)]`` (a lone 0/1 vector of that length
        is read as bits, like the reference's ``achieved_goal_vector[0]``)."""
        g = torch.as_tensor(g, device=self.device)
        nb = len(self.task_list)
        is_bits = (g.dim() >= 2 and g.shape[-1] == nb) or (g.dim() == 1 and g.numel() == nb and int(g.max()) <= 1)
        if is_bits:
            return (g.reshape(-1, nb).to(torch.int32) << self._bits).sum(dim=1).to(torch.int32)
        return g.to(torch.int32).reshape(-1)

    def compute_reward_equal(self, achieved_goal=None, desired_goal=None, info=None):
        """``compute_reward_equal`` (``ray.py:757-761``): MAX_STEPS where achieved == desired else -1."""
        a, d = self._to_mask(achieved_goal), self._to_mask(desired_goal)
        return torch.where(a == d, self.MAX_STEPS, -1).to(torch.int32)

    def compute_reward_subset(self, achieved_goal=None, desired_goal=None, info=None):
        """``compute_reward_subset`` (``ray.py:763-767``): MAX_STEPS where desired is a subset of achieved."""
        a, d = self._to_mask(achieved_goal), self._to_mask(desired_goal)
        return torch.where((d & ~a) == 0, self.MAX_STEPS, -1).to(torch.int32)

    # ---- state injection / extraction (parity tests upload reference-generated states) ---------------------
    def load_state(self, grid, r, c, hold, desired, achieved=None, t=None, init_grid=None):
        """Inject compact states: ``grid uint8[N,H,W]``, ``r,c,hold,desired[N]``."""
        N, H, W, dev = self.num_envs, self.cfg.H, self.cfg.W, self.device

        def vec(x, dtype=torch.int32):
            return torch.as_tensor(np.asarray(x), device=dev).to(dtype).reshape(N)

        g = torch.as_tensor(np.asarray(grid), device=dev).to(torch.uint8).reshape(N, H * W)
        self.grid.zero_()
        self.grid[:, :H * W] = g
        if init_grid is None:
            self.init_grid.copy_(self.grid)
        else:
            self.init_grid.zero_()
            self.init_grid[:, :H * W] = torch.as_tensor(np.asarray(init_grid), device=dev).to(torch.uint8).reshape(N, H * W)
        self.agent.copy_(vec(r) | (vec(c) << 8) | (vec(hold) << 16))
        ach = torch.zeros(N, dtype=torch.int32, device=dev) if achieved is None else vec(achieved)
        self.goal.copy_(ach | (vec(desired) << 16))
        if t is None:
            self.t.zero_()
        else:
            self.t.copy_(vec(t))
        self.init_agent.copy_(self.agent)
        if self.obs_mode == "pixels":
            self.render()
            if self.goal_images:
                self.init_obs.copy_(self.obs)
                with self._guard():
                    _lib.check(self._lib.cw_imagine(C.byref(self.cfg), C.byref(self._state), self.desired_goal.data_ptr(),
                                                    self._stream()), "cw_imagine")
        elif self.obs_mode == "onehot":
            with self._guard():
                _lib.check(self._lib.cw_imagine(C.byref(self.cfg), C.byref(self._state), None, self._stream()), "cw_imagine")
        self._is_reset = True
        self._obs_version += 1
        return self._observation()

    def export_state(self):
        """Compact state as NumPy arrays (Appendix B.3)."""
        H, W = self.cfg.H, self.cfg.W
        ag, gl = self.agent.cpu().numpy().astype(np.uint32), self.goal.cpu().numpy().astype(np.uint32)
        return {"grid": self.grid[:, :H * W].reshape(-1, H, W).cpu().numpy(),
                "init_grid": self.init_grid[:, :H * W].reshape(-1, H, W).cpu().numpy(),
                "r": (ag & 0xFF).astype(np.uint8), "c": ((ag >> 8) & 0xFF).astype(np.uint8),
                "hold": ((ag >> 16) & 0xFF).astype(np.uint8), "achieved": (gl & 0xFFFF).astype(np.uint16),
                "desired": (gl >> 16).astype(np.uint16), "t": self.t.cpu().numpy(), "episode": self.episode.cpu().numpy()}

    # ---- statistics ------------------------------------------------------------------------------------------
    def episode_stats(self, stats=None):
        """Finished-episode statistics accumulated on the device (this rank's worlds unless ``stats`` is given)."""
        s = (self.stats if stats is None else stats).cpu().numpy()
        n = max(int(s[0]), 1)
        return {"episodes": int(s[0]), "successes": int(s[1]), "success_rate": s[1] / n, "mean_return": s[2] / n,
                "mean_length": s[3] / n, "achieved_counts": s[4:13].tolist(), "desired_counts": s[13:22].tolist()}

    def close(self):
        pass


class BatchedCraftingWorldEnvFlat(BatchedCraftingWorldEnv):
    """Batched mirror of ``CraftingWorldEnvFlat`` (``craftingworld_flat.py:40-57, 119, 185``): identical dynamics,
    8x8 grid / 100 steps by default, no ``fixed_init_state``, and ``reset``/``step`` return the bare image tensor."""

    def __init__(self, num_envs, size=(8, 8), max_steps=100, store_gif=False, render_save_rate=1, task_list=TASK_LIST,
                 selected_tasks=TASK_LIST, number_of_tasks=None, stacking=True, reward_style=None, **kw):
        kw.setdefault("goal_images", False)
        super().__init__(num_envs, size=size, fixed_init_state=0, max_steps=max_steps, store_gif=store_gif,
                         render_save_rate=render_save_rate, task_list=task_list, selected_tasks=selected_tasks,
                         number_of_tasks=number_of_tasks, stacking=stacking, reward_style=reward_style, obs_mode="pixels", **kw)
        self.observation_space = spaces.Box(0, 255, (4 * self.cfg.W, 4 * self.cfg.H, 3), np.uint8)   # flat.py:57

    def reset(self, mask=None):
        return super().reset(mask)["observation"]                                          # flat.py:119

    def step(self, actions):
        obs, reward, done, info = super().step(actions)
        return obs["observation"], reward, done, info                                      # flat.py:185


class BatchedCraftingWorldEnvOneHot(BatchedCraftingWorldEnv):
    """Batched mirror of ``CraftingWorldEnvOneHot`` (``carftingworld_onehot.py``): the observation is the one-hot state
    ``uint8[N, H, W, 12]`` (no rendering), ``desired_goal`` the imagined one-hot state (``:310``)."""

    def __init__(self, num_envs, *args, **kw):
        kw["obs_mode"] = "onehot"
        super().__init__(num_envs, *args, **kw)
        vec = self.observation_vector_space["observation"]
        self.observation_space = spaces.Dict(dict(observation=vec, desired_goal=vec, achieved_goal=vec,
                                                  init_observation=vec))                   # onehot.py:84-103


class AltObs(Mapping):
    """Observation dict of the AltObs variant: int16 frames rendered lazily from the compact device state."""
    _KEYS = ("observation", "desired_goal", "achieved_goal", "init_observation")

    def __init__(self, env):
        self._env, self._version, self._cache = env, env._obs_version, {}

    def __getitem__(self, k):
        env = self._env
        if self._version != env._obs_version:
            self._version, self._cache = env._obs_version, {}
        key = "observation" if k == "achieved_goal" else k
        if key not in self._KEYS:
            raise KeyError(k)
        if key not in self._cache:
            if key == "observation":
                self._cache[key] = env.render_alt(env.grid, env.agent)
            elif key == "desired_goal":
                self._cache[key] = env.render_alt(env.goal_grid, env.goal_agent)
            else:
                self._cache[key] = env.render_alt(env.init_grid, env.init_agent)
        return self._cache[key]

    def __iter__(self):
        return iter(self._KEYS)

    def __len__(self):
        return len(self._KEYS)


class BatchedCraftingWorldEnvAltObs(BatchedCraftingWorldEnv):
    """Batched mirror of ``CraftingWorldEnvAltObs`` (``craftingworld_altobs.py``; unregistered upstream): the dynamics
    of ``CraftingWorldEnvRay`` with the 3x3-sub-pixel renderer (``:489-548``).  Frames are ``int16[N, 3H+3, 3W, 3]``
    because pixel values reach 510 upstream; ``stacked_obs=True`` returns the four frames stacked on axis 1
    (``:116-119, 258-259, 408-410``)."""

    def __init__(self, num_envs, *args, stacked_obs=False, **kw):
        kw["obs_mode"] = "onehot"                      # compact goal / init state; frames are rendered from it
        super().__init__(num_envs, *args, **kw)
        self.stacked_obs = stacked_obs is True
        pw, ph = (self.cfg.W + 1) * 3, self.cfg.H * 3                                      # altobs.py:115
        img = spaces.Box(0, 255, (pw, ph, 3), np.int16)
        self.observation_space = (spaces.Box(0, 255, (4, pw, ph, 3), np.int16) if self.stacked_obs else
                                  spaces.Dict(dict(observation=img, desired_goal=img, achieved_goal=img, init_observation=img)))

    def render_alt(self, grid, agent):
        M = grid.shape[0]
        out = torch.empty((M, 3 * self.cfg.H + 3, 3 * self.cfg.W, 3), dtype=torch.int16, device=self.device)
        with self._guard():
            _lib.check(self._lib.cw_render_alt(C.byref(self.cfg), grid.data_ptr(), agent.data_ptr(), out.data_ptr(), M,
                                               self._stream()), "cw_render_alt")
        return out

    def _observation(self):
        obs = AltObs(self)
        if self.stacked_obs:
            return torch.stack([obs["observation"], obs["desired_goal"], obs["achieved_goal"], obs["init_observation"]], dim=1)
        return obs

    def render(self, state=None, mode="Non", tile_size=4):
        """``render`` (``craftingworld_altobs.py:489-548``): the current worlds, or foreign states given as
        ``(grid uint8[M,H,W], r[M], c[M], hold[M])`` -> a fresh ``int16[M, 3H+3, 3W, 3]`` tensor."""
        if state is None:
            return self.render_alt(self.grid, self.agent)
        H, W = self.cfg.H, self.cfg.W
        g, r, c, h = state
        g = torch.as_tensor(g, device=self.device).to(torch.uint8).reshape(-1, H * W)
        grid = torch.zeros((g.shape[0], self.cfg.cell_stride), dtype=torch.uint8, device=self.device)
        grid[:, :H * W] = g
        r, c, h = (torch.as_tensor(x, device=self.device).to(torch.int32).reshape(-1) for x in (r, c, h))
        return self.render_alt(grid, (r | (c << 8) | (h << 16)).contiguous())
