"""Ecosystem adapters: a VectorEnv-style facade, registration under gym/gymnasium when one of them is
importable, and a host-side episode GIF recorder fed from device frames.

Neither ``gym`` nor ``gymnasium`` is a dependency (they are not installable in the target image); the facade follows the
``gymnasium.vector.VectorEnv`` call shapes (``reset -> (obs, info)``, ``step -> (obs, reward, terminated, truncated,
info)``) so it can be handed to libraries that only duck-type the vector API.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import spaces
from .env import BatchedCraftingWorldEnv


class CraftingWorldVectorEnv:
    """``gymnasium.vector``-shaped view of a :class:`BatchedCraftingWorldEnv` (auto-reset, same-step semantics:
    the observation returned with ``terminated | truncated`` is the first of the next episode)."""

    def __init__(self, num_envs, to_numpy=False, **kw):
        kw.setdefault("auto_reset", True)
        self.env = BatchedCraftingWorldEnv(num_envs, **kw)
        self.num_envs, self.to_numpy = self.env.num_envs, bool(to_numpy)
        self.single_observation_space = self.env.observation_space
        self.single_action_space = self.env.action_space
        self.observation_space = self.env.observation_space          # per-world spaces; batch axis is num_envs
        self.action_space = spaces.Box(0, 5, (self.num_envs,), np.int64)
        self.is_vector_env = True
        self.closed = False

    def _out(self, x):
        if isinstance(x, torch.Tensor) and self.to_numpy:
            return x.cpu().numpy()
        return x

    def _obs(self, obs):
        return {k: self._out(v) for k, v in obs.items()} if hasattr(obs, "items") else self._out(obs)

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.env.seed(seed)
        return self._obs(self.env.reset()), {}

    def step(self, actions):
        obs, reward, done, info = self.env.step(actions)
        terminated = reward == self.env.MAX_STEPS            # task completed (ray.py:367: reward == MAX_STEPS)
        truncated = done & ~terminated                       # step_num >= MAX_STEPS
        return (self._obs(obs), self._out(reward), self._out(terminated), self._out(truncated),
                {"achieved_mask": self._out(self.env.achieved_mask), "desired_mask": self._out(self.env.desired_mask)})

    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        return self.step(self._pending)

    def close(self):
        self.closed = True


# the reference's three registrations (gym_craftingworld/__init__.py:5-18): id -> (its entry-point class, our batched mirror)
REGISTRATIONS = {
    "craftingworld-v3": ("CraftingWorldEnvRay", "BatchedCraftingWorldEnv"),
    "craftingworldflat-v3": ("CraftingWorldEnvFlat", "BatchedCraftingWorldEnvFlat"),
    "craftingworldonehot-v3": ("CraftingWorldEnvOneHot", "BatchedCraftingWorldEnvOneHot"),
}


def register_envs(num_envs=4096, reference_ids=False, register=None):
    """Register the three batched mirrors with gymnasium or gym when one of them is importable -- what
    ``gym_craftingworld/__init__.py:5-18`` does for the reference: the same three environments with the same ``kwargs``
    (``stacking=True, render_save_rate=10``) plus ``num_envs``.  By default the ids carry a ``-b200`` tag
    (``craftingworld-b200-v3``, ``craftingworldflat-b200-v3``, ``craftingworldonehot-b200-v3``) so that both packages can
    be installed side by side; ``reference_ids=True`` registers the reference's own ids instead (a drop-in for code that
    calls ``gym.make('craftingworld-v3')``).  ``register`` overrides the registration function (tests).
    Returns the ids registered (empty when neither package exists)."""
    if register is None:
        for mod in ("gymnasium", "gym"):
            try:
                register = __import__(mod + ".envs.registration", fromlist=["register"]).register
                break
            except Exception:  # noqa: BLE001
                continue
    if register is None:
        return []
    ids = []
    for ref_id, (_, cls) in REGISTRATIONS.items():
        env_id = ref_id if reference_ids else ref_id.replace("-v3", "-b200-v3")
        register(id=env_id, entry_point=f"gym_craftingworld_b200:{cls}",
                 kwargs={"num_envs": num_envs, "stacking": True, "render_save_rate": 10})
        ids.append(env_id)
    return ids


class GifRecorder:
    """Host-side episode recorder (the role of ``allow_gif_storage`` / ``__render_gif``, ``ray.py:565-597, 769-782``):
    call :meth:`capture` after every ``reset``/``step``; the frames of world ``index`` are pulled from the device and an
    animated GIF is written to ``renders/env{env_id}/`` whenever that world finishes an episode."""

    def __init__(self, index=0, directory="renders", env_id=0, scale=4):
        self.index, self.scale = int(index), int(scale)
        self.dir = os.path.join(directory, f"env{env_id}")
        os.makedirs(self.dir, exist_ok=True)
        self.frames, self.episode, self.saved = [], 0, []

    def capture(self, obs, done=None):
        frame = obs["observation"] if hasattr(obs, "keys") else obs
        frame = frame[self.index]
        frame = frame.cpu().numpy() if isinstance(frame, torch.Tensor) else np.asarray(frame)
        finished = bool(done[self.index]) if done is not None else False
        if finished and self.frames:
            self.save()
        self.frames.append(np.clip(frame, 0, 255).astype(np.uint8))

    def save(self):
        from PIL import Image
        imgs = [Image.fromarray(f).resize((f.shape[1] * self.scale, f.shape[0] * self.scale), Image.NEAREST) for f in self.frames]
        path = os.path.join(self.dir, f"E{self.episode}({len(imgs) - 1}).gif")
        imgs[0].save(path, save_all=True, append_images=imgs[1:], duration=100, loop=0)
        self.saved.append(path)
        self.frames, self.episode = [], self.episode + 1
        return path
