#!/usr/bin/env python
"""Generate the frozen golden traces in this directory by driving the UNMODIFIED reference env
(``/root/reference``, imported under ``oracle/ref_shim.py``).  Run in the builder container only:

    python tests/golden/make_golden.py

The reference has no golden vectors of its own for this path, so these outputs of the
reference itself are the parity pin for ``oracle/`` and, through it and directly, for the CUDA kernels.

File format (``*.npz``, one batch of B independent worlds stepped T times with NO reset in between --
stepping past ``done`` is legal upstream, ``ray.py:367``):
  H, W, max_steps, subset           scalars (subset=1 -> reward_style set -> compute_reward_subset)
  grid0 u8[B,H,W], r0 c0 hold0 u8[B], desired u16[B]      injected initial state (Appendix B.2)
  actions u8[B,T]
  grid u8[B,T,H,W], r c hold u8[B,T], achieved u16[B,T], reward i32[B,T], done u8[B,T]   state AFTER step t
  frame_crc u32[B,T]                crc32 of the uint8 frame after step t (every step)
  frame_t i32[F], frames u8[B,F,4H,4W,3]   full frames at the listed steps; frame0 u8[B,4H,4W,3] at reset
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402

ray = ref_shim.load_reference()


def crc(img):
    return zlib.crc32(np.ascontiguousarray(img, dtype=np.uint8).tobytes()) & 0xFFFFFFFF


def run_batch(path, H, W, worlds, actions, max_steps, subset, frame_t):
    """worlds: list of (grid, r, c, hold, desired)."""
    B, T = actions.shape
    out = dict(H=H, W=W, max_steps=max_steps, subset=int(subset),
               grid0=np.zeros((B, H, W), np.uint8), r0=np.zeros(B, np.uint8), c0=np.zeros(B, np.uint8),
               hold0=np.zeros(B, np.uint8), desired=np.zeros(B, np.uint16), actions=actions.astype(np.uint8),
               grid=np.zeros((B, T, H, W), np.uint8), r=np.zeros((B, T), np.uint8), c=np.zeros((B, T), np.uint8),
               hold=np.zeros((B, T), np.uint8), achieved=np.zeros((B, T), np.uint16),
               reward=np.zeros((B, T), np.int32), done=np.zeros((B, T), np.uint8),
               frame_crc=np.zeros((B, T), np.uint32), frame_t=np.asarray(frame_t, np.int32),
               frames=np.zeros((B, len(frame_t), 4 * H, 4 * W, 3), np.uint8),
               frame0=np.zeros((B, 4 * H, 4 * W, 3), np.uint8))
    fidx = {int(t): i for i, t in enumerate(frame_t)}
    for b, (grid, r, c, hold, desired) in enumerate(worlds):
        env = ref_shim.make_injected_env(ray, grid, r, c, hold, desired, max_steps=max_steps,
                                         reward_style=("subset" if subset else None))
        out["grid0"][b], out["r0"][b], out["c0"][b], out["hold0"][b], out["desired"][b] = grid, r, c, hold, desired
        out["frame0"][b] = env.obs_image.astype(np.uint8)
        for t in range(T):
            obs, reward, done, info = env.step(int(actions[b, t]))
            g, rr, cc, hh, ach, px = ref_shim.read_back(env)
            assert obs["observation"] is env.obs_image
            out["grid"][b, t], out["r"][b, t], out["c"][b, t], out["hold"][b, t] = g, rr, cc, hh
            out["achieved"][b, t], out["reward"][b, t], out["done"][b, t] = ach, reward, done
            out["frame_crc"][b, t] = crc(px)
            if t in fidx:
                out["frames"][b, fidx[t]] = px
    np.savez_compressed(path, **out)
    print(f"{os.path.basename(path)}: B={B} T={T} {H}x{W}  successes={(out['reward'] > 0).sum()} "
          f"size={os.path.getsize(path) / 1024:.0f} KiB")


def dense_world(rng, H, W, density, hold=None):
    """Dense synthetic placement (BASELINE config 5 style): each cell occupied w.p. density, type uniform over
    the 8 objects, >=1 of each type when room allows; agent anywhere (possibly ON an object); random held item."""
    grid = np.where(rng.random_sample((H, W)) < density, rng.randint(1, 9, (H, W)), 0).astype(np.uint8)
    cells = rng.permutation(H * W)[:8]
    if H * W >= 12:
        for k, cell in enumerate(cells):
            grid[cell // W, cell % W] = k + 1
    r, c = int(rng.randint(H)), int(rng.randint(W))
    if rng.random_sample() < 0.7:
        grid[r, c] = 0
    if hold is None:
        hold = int(rng.choice([0, 0, 1, 2, 3]))
    desired = int(rng.randint(1, 512))
    if rng.random_sample() < 0.35:                      # easy goals so that successes occur
        desired = int(1 << rng.choice([1, 3, 4, 5]))
    return grid, r, c, hold, desired


def gen_dense(name, H, W, B, T, seed, max_steps, subset, frame_stride):
    rng = np.random.RandomState(seed)
    worlds = [dense_world(rng, H, W, rng.choice([0.15, 0.3, 0.5, 0.8])) for _ in range(B)]
    # action mix: mostly uniform; some envs biased to pickup/drop so the held-item state machine is exercised
    actions = rng.randint(0, 6, (B, T))
    frame_t = list(range(0, T, frame_stride))
    run_batch(os.path.join(HERE, name), H, W, worlds, actions, max_steps, subset, frame_t)


def gen_sampled(name, H, W, B, T, seed, max_steps, frame_stride):
    """Worlds produced by the reference's own reset() (sample_state + task sampling, ray.py:156-218)."""
    rng = np.random.RandomState(seed)
    worlds = []
    for b in range(B):
        env = ray.CraftingWorldEnvRay(size=(W, H), max_steps=max_steps)
        env.seed(seed * 1000 + b)
        env.reset()
        g, r, c, h, _, _ = ref_shim.read_back(env)
        worlds.append((g, r, c, h, ref_shim.bits_to_mask(env.desired_goal_vector[0])))
    actions = rng.randint(0, 6, (B, T))
    run_batch(os.path.join(HERE, name), H, W, worlds, actions, max_steps, False, list(range(0, T, frame_stride)))


def gen_cfg1(name):
    """BASELINE config 1: default grid, task [ChopTree, BuildHouse] forced, RandomState(0) actions, 1000 steps,
    reset() on done.  Episodes are stored as separate worlds of one batch (padded with action 6 = no-op never
    sent to the reference: rows past the episode end are marked by len)."""
    env = ray.CraftingWorldEnvRay(size=(21, 21), selected_tasks=['ChopTree', 'BuildHouse'], number_of_tasks=2)
    env.seed(0)
    acts = np.random.RandomState(0).randint(0, 6, 1000)
    episodes, cur = [], None
    need_reset = True
    for t in range(1000):
        if need_reset:
            env.reset()
            env.desired_goal_vector[0, [2, 3]] = 1
            g, r, c, h, _, _ = ref_shim.read_back(env)
            cur = dict(world=(g, r, c, h, ref_shim.bits_to_mask(env.desired_goal_vector[0])), actions=[])
            episodes.append(cur)
        cur["actions"].append(int(acts[t]))
        _, _, need_reset, _ = env.step(int(acts[t]))
    T = max(len(e["actions"]) for e in episodes)
    # replay each episode through the injected-state path so the file has the common layout
    actions = np.zeros((len(episodes), T), np.int64)
    for i, e in enumerate(episodes):
        a = e["actions"] + [4] * (T - len(e["actions"]))     # pad with pickup (agent on empty cell: harmless)
        actions[i] = a
    run_batch(os.path.join(HERE, name), 21, 21, [e["world"] for e in episodes], actions, 300, False,
              list(range(0, T, 25)))


S, A, Hm, R, Tr, Br, Ho, Wh = 1, 2, 3, 4, 5, 6, 7, 8


def gen_quirks(name):
    """Hand-built known-answer scenarios, one per Appendix-C quirk, answered by the reference."""
    H = W = 5
    worlds, scripts = [], []

    def world(cells, r, c, hold, desired):
        g = np.zeros((H, W), np.uint8)
        for (rr, cc), v in cells.items():
            g[rr, cc] = v
        return g, r, c, hold, desired

    # 0: chop tree with axe, then build house with hammer (full craft chain); desired ChopTree|BuildHouse|GoToHouse
    worlds.append(world({(0, 1): A, (0, 3): Tr, (2, 0): Hm}, 0, 0, 0, (1 << 3) | (1 << 2) | (1 << 5)))
    scripts.append([1, 4, 1, 1, 3, 5, 3, 3, 2, 2, 4, 0, 0, 1, 1, 1])
    # 1: tree / rock block without tool; failed moves still evaluate tasks; reward -1 on unchanged state
    worlds.append(world({(0, 1): Tr, (1, 0): R, (2, 2): Ho}, 0, 0, 0, 1 << 5))
    scripts.append([1, 2, 0, 3, 1, 2, 4, 5, 1, 2, 0, 3, 1, 2, 4, 5])
    # 2: wheat -> bread with axe, bread not eaten until re-entry; EatBread on re-entry
    worlds.append(world({(1, 1): Wh, (0, 0): A}, 0, 0, 0, (1 << 0) | (1 << 1)))
    scripts.append([4, 2, 1, 1, 3, 0, 2, 5, 4, 0, 2, 3, 1, 1, 3, 3])
    # 3: GoToHouse is level triggered; walking on/off the house; goal == GoToHouse succeeds immediately
    worlds.append(world({(0, 2): Ho}, 0, 0, 0, (1 << 5) | (1 << 4)))
    scripts.append([1, 1, 1, 3, 3, 1, 1, 2, 0, 0, 3, 1, 4, 5, 1, 3])
    # 4: MoveSticks judged while carrying vs INITIAL grid; back over the initial cell flips it to 0
    worlds.append(world({(0, 1): S}, 0, 0, 0, 1 << 8))
    scripts.append([1, 4, 1, 3, 3, 1, 5, 1, 3, 4, 2, 0, 5, 1, 3, 2])
    # 5: MoveSticks with tree-home rule: chop tree, pick up the resulting sticks, carry them back
    worlds.append(world({(0, 1): A, (0, 3): Tr}, 0, 0, 0, (1 << 8) | (1 << 3)))
    scripts.append([1, 4, 1, 1, 3, 5, 1, 4, 1, 3, 3, 1, 1, 5, 2, 0])
    # 6: hammer breaks rock by walking; MoveHammer; drop on occupied cell fails; pickup while holding fails
    worlds.append(world({(0, 1): Hm, (0, 2): R, (1, 1): A}, 0, 0, 0, (1 << 4) | (1 << 7)))
    scripts.append([1, 4, 1, 1, 2, 3, 5, 4, 3, 1, 5, 4, 0, 2, 5, 5])
    # 7: edge clamps in all four corners
    worlds.append(world({(4, 4): Br}, 0, 0, 0, 1 << 1))
    scripts.append([0, 3, 2, 2, 2, 2, 2, 3, 1, 1, 1, 1, 1, 2, 0, 3])
    # 8: success then keep stepping past done; extra sticky skill makes equal-style unwinnable
    worlds.append(world({(0, 1): Br, (0, 2): R, (1, 0): Hm}, 0, 0, 0, 1 << 1))
    scripts.append([1, 1, 3, 3, 2, 4, 0, 1, 1, 1, 3, 3, 2, 2, 0, 0])
    # 9: start already holding (injected) and standing on an object
    worlds.append(world({(2, 2): S, (2, 3): Wh, (1, 2): Tr}, 2, 2, 3, (1 << 2) | (1 << 7)))
    scripts.append([4, 5, 1, 3, 0, 2, 2, 0, 3, 1, 1, 1, 4, 5, 0, 0])
    for subset, nm in ((False, name), (True, name.replace(".npz", "_subset.npz"))):
        run_batch(os.path.join(HERE, nm), H, W, worlds, np.asarray(scripts), 12, subset, list(range(16)))


def gen_altobs(name, src="dense_8x8.npz", stride=6):
    """AltObs frames (craftingworld_altobs.py render, int values up to 510) of the states along an existing trace:
    the reference AltObs env is injected with each state and asked for render(state)."""
    alt = ref_shim.load_reference_altobs()
    d = np.load(os.path.join(HERE, src))
    H, W = int(d["H"]), int(d["W"])
    env = alt.CraftingWorldEnvAltObs(size=(W, H))
    B, T = d["actions"].shape
    ts = list(range(0, T, stride))
    frames = np.zeros((B, len(ts), 3 * H + 3, 3 * W, 3), np.int16)
    for b in range(B):
        for i, t in enumerate(ts):
            st = ref_shim.compact_to_onehot(d["grid"][b, t], int(d["r"][b, t]), int(d["c"][b, t]), int(d["hold"][b, t]))
            frames[b, i] = env.render(st)
    np.savez_compressed(os.path.join(HERE, name), src=src, frame_t=np.asarray(ts, np.int32), frames=frames)
    print(f"{name}: {frames.shape} max pixel {frames.max()}")


if __name__ == "__main__":
    if "--altobs-only" in sys.argv:
        gen_altobs("altobs_8x8.npz")
        sys.exit(0)
    gen_quirks("quirks_5x5.npz")
    gen_cfg1("cfg1_21x21.npz")
    gen_sampled("sampled_21x21.npz", 21, 21, 48, 320, 7, 300, 40)
    gen_dense("dense_4x4.npz", 4, 4, 96, 48, 11, 40, False, 1)
    gen_dense("dense_5x5_subset.npz", 5, 5, 96, 64, 12, 50, True, 1)
    gen_dense("dense_8x8.npz", 8, 8, 96, 96, 13, 80, False, 4)
    gen_dense("dense_21x21.npz", 21, 21, 64, 160, 14, 120, False, 32)
    gen_dense("dense_32x32.npz", 32, 32, 32, 160, 15, 150, False, 40)
    gen_altobs("altobs_8x8.npz")
