"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- a single-world Python + NumPy env with the reference's cost model.

The reference (``CraftingWorldEnvRay``) cannot travel to the GPU box (``gym`` / ``matplotlib`` are not installable
and ``/root/reference`` is absent there), so ``bench.py --impl reference`` and the ``cpu_baseline`` leg time this
port instead: one Python object per world, NumPy RandomState reset (permutation placement, ``ray.py:605-613``),
the step logic of ``oracle/compact.py`` and -- like the reference -- an INCREMENTAL re-render of the <= 2 changed
cells per step (``render_edit``, ``ray.py:522-557``) with a full ``render`` only at reset (``ray.py:192``).
It runs on the compact state, so it does fewer NumPy calls per step than the reference does on its one-hot
``int64[H,W,12]`` state: as a baseline it is conservative (faster than the real reference).
"""
from __future__ import annotations

import numpy as np

from . import compact


class PortEnv:
    def __init__(self, size=(21, 21), max_steps=300, selected=tuple(range(9)), number_of_tasks=None, stacking=True,
                 subset_reward=False, seed=None):
        W, H = size
        self.cfg = compact.Config(H=H, W=W, max_steps=max_steps, subset_reward=subset_reward, stacking=stacking,
                                  selected=tuple(selected),
                                  number_of_tasks=number_of_tasks if number_of_tasks is not None else len(selected))
        self.np_random = np.random.RandomState(seed)
        self.s = None
        self.obs_image = None

    def reset(self):
        cfg, rng = self.cfg, self.np_random
        n = rng.randint(cfg.number_of_tasks) + 1 if cfg.stacking else 1              # ray.py:169
        idx = np.arange(len(cfg.selected))
        rng.shuffle(idx)                                                             # ray.py:171-172
        desired = 0
        for i in idx[:n]:
            desired |= 1 << cfg.selected[i]                                          # ray.py:173-174
        perm = np.arange(cfg.H * cfg.W)
        rng.shuffle(perm)                                                            # ray.py:610-611
        codes = np.zeros(cfg.H * cfg.W, np.uint8)
        codes[:9] = np.arange(1, 10)
        flat = codes[perm]                                                           # ray.py:612
        agent = int(np.flatnonzero(flat == 9)[0])
        flat[agent] = 0
        grid = flat.reshape(cfg.H, cfg.W)
        self.s = compact.EnvState(grid, grid.copy(), agent // cfg.W, agent % cfg.W, 0, 0, desired, 0)
        self.obs_image = compact.render(grid, self.s.r, self.s.c, 0)                 # ray.py:192
        return self.obs_image

    def _render_cell(self, r, c):                                                    # ray.py:550-557
        s, img = self.s, self.obs_image
        img[4 * r:4 * r + 4, 4 * c:4 * c + 4] = compact.LUT[s.grid[r, c]]
        if (r, c) == (s.r, s.c):
            img[4 * r + 1:4 * r + 3, 4 * c + 1:4 * c + 3] = 255
            if s.hold:
                img[4 * r + 2, 4 * c + 1:4 * c + 3] = compact.LUT[s.hold]

    def step(self, a):
        s = self.s
        r0, c0 = s.r, s.c
        reward, done, changed = compact.step_env(s, a, self.cfg)
        if changed:                                                                  # ray.py:348-358
            self._render_cell(r0, c0)
            if (s.r, s.c) != (r0, c0):
                self._render_cell(s.r, s.c)
        return self.obs_image, reward, done, None


def run_worker(args):
    """One process of the CPU-baseline arm: ``envs`` worlds, ``warmup`` + ``steps`` batched steps (each steps every
    world once, reset on done).  Returns (env_steps, seconds) for the timed part."""
    import time
    envs, steps, warmup, size, max_steps, seed = args
    rng = np.random.RandomState(seed)
    worlds = [PortEnv(size=size, max_steps=max_steps, seed=seed * 7919 + i) for i in range(envs)]
    for w in worlds:
        w.reset()
    acts = rng.randint(0, 6, (warmup + steps, envs))
    t0 = 0.0
    for k in range(warmup + steps):
        if k == warmup:
            t0 = time.perf_counter()
        for i, w in enumerate(worlds):
            _, _, done, _ = w.step(int(acts[k, i]))
            if done:
                w.reset()
    return envs * steps, time.perf_counter() - t0


def run_worker_reference(args):
    """The same loop as :func:`run_worker` on the UNMODIFIED reference class ``CraftingWorldEnvRay``
    (``ray.py:53-378``; ``/root/reference`` or its install under ``oracle/_ref``, see ``oracle/build_ref.py``): one
    reference env object per world, ``env.step(a)`` then ``env.reset()`` on done (``gen_info.rst:71-80``).
    Returns (env_steps, seconds) for the timed part."""
    import time
    from . import ref_shim
    envs, steps, warmup, size, max_steps, seed = args
    ray = ref_shim.load_reference()
    rng = np.random.RandomState(seed)
    worlds = []
    for i in range(envs):
        w = ray.CraftingWorldEnvRay(size=size, max_steps=max_steps)                  # nine-skill random tasks, stacking
        w.seed(seed * 7919 + i)
        w.reset()
        worlds.append(w)
    acts = rng.randint(0, 6, (warmup + steps, envs))
    t0 = 0.0
    for k in range(warmup + steps):
        if k == warmup:
            t0 = time.perf_counter()
        for i, w in enumerate(worlds):
            _, _, done, _ = w.step(int(acts[k, i]))
            if done:
                w.reset()
    return envs * steps, time.perf_counter() - t0
