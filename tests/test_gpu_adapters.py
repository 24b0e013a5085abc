"""GPU (-m gpu): the observation-format adapters of SURVEY 8(f)-2 -- batched mirrors of CraftingWorldEnvFlat
(craftingworld_flat.py) and CraftingWorldEnvOneHot (carftingworld_onehot.py) -- against golden traces / the oracle."""
import numpy as np
import pytest
import torch

from oracle import compact, native, ref_shim
from tests import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cw():
    import gym_craftingworld_b200 as pkg
    assert torch.cuda.is_available()
    return pkg


def test_flat_env_defaults_and_trace(cw):
    env = cw.BatchedCraftingWorldEnvFlat(4, seed=0)
    assert (env.STATE_W, env.STATE_H, env.MAX_STEPS) == (8, 8, 100)            # craftingworld_flat.py:40-43
    assert env.observation_space.shape == (32, 32, 3)
    obs = env.reset()
    assert isinstance(obs, torch.Tensor) and obs.shape == (4, 32, 32, 3)       # bare image, flat.py:119
    d = gu.load("dense_8x8.npz")
    B, T = d["actions"].shape
    env = cw.BatchedCraftingWorldEnvFlat(B, size=(8, 8), max_steps=d["max_steps"], seed=0, auto_reset=False)
    env.load_state(d["grid0"], d["r0"], d["c0"], d["hold0"], d["desired"])
    for t in range(T):
        obs, reward, done, info = env.step(d["actions"][:, t])
        assert isinstance(obs, torch.Tensor)
        assert np.array_equal(reward.cpu().numpy(), d["reward"][:, t]) and np.array_equal(done.cpu().numpy(), d["done"][:, t].astype(bool))
        assert [gu.crc(f) for f in obs.cpu().numpy()] == list(d["frame_crc"][:, t]), t
    assert set(info) == {"task_success", "desired_goal", "achieved_goal"}


def test_onehot_env_matches_oracle(cw):
    N, size, seed, K = 96, 6, 31, 40
    env = cw.BatchedCraftingWorldEnvOneHot(N, size=(size, size), max_steps=9, seed=seed)
    assert env.observation_space["observation"].shape == (size, size, 12)     # carftingworld_onehot.py:84-103
    cfg = native.make_config(H=size, W=size, max_steps=9)
    ob = native.OracleBatch(cfg, N, seed=seed)
    pc = compact.Config(H=size, W=size, max_steps=9)

    def onehot_of(grid, agent):
        return np.stack([ref_shim.compact_to_onehot(grid[n, :size * size].reshape(size, size), int(agent[n] & 0xFF),
                                                    int((agent[n] >> 8) & 0xFF), int((agent[n] >> 16) & 0xFF)) for n in range(N)]).astype(np.uint8)

    def goal_of(episodes):
        out = []
        for n in range(N):
            _, (g, r, c, h) = compact.reset_env(seed, n, int(episodes[n]) - 1, pc, with_goal=True)
            out.append(ref_shim.compact_to_onehot(g, r, c, h))
        return np.stack(out).astype(np.uint8)

    obs = env.reset()
    ob.reset()
    assert obs["achieved_goal"] is obs["observation"]
    assert np.array_equal(obs["observation"].cpu().numpy(), onehot_of(ob.grid, ob.agent))
    assert np.array_equal(obs["init_observation"].cpu().numpy(), onehot_of(ob.grid, ob.agent))   # INIT_OBS copy, onehot.py:203
    assert np.array_equal(obs["desired_goal"].cpu().numpy(), goal_of(ob.episode))             # imagined state, onehot.py:310
    rng = np.random.RandomState(3)
    init_grid, init_agent = ob.grid.copy(), ob.agent.copy()
    for k in range(K):
        a = rng.randint(0, 6, N).astype(np.uint8)
        obs, reward, done, _ = env.step(a)
        o_reward, o_done = ob.step_full(a, auto_reset=True)
        fresh = o_done == 1
        init_grid[fresh], init_agent[fresh] = ob.grid[fresh], ob.agent[fresh]
        assert np.array_equal(reward.cpu().numpy(), o_reward) and np.array_equal(done.cpu().numpy(), fresh)
        assert np.array_equal(obs["observation"].cpu().numpy(), onehot_of(ob.grid, ob.agent)), k
        if k % 8 == 0 or k == K - 1:
            assert np.array_equal(obs["desired_goal"].cpu().numpy(), goal_of(ob.episode)), k
            assert np.array_equal(obs["init_observation"].cpu().numpy(), onehot_of(init_grid, init_agent)), k
    assert ob.stats[0] > 0 and np.array_equal(env.stats.cpu().numpy(), ob.stats)


def test_onehot_goal_state_renders_to_goal_frame(cw):
    """The compact goal state of the one-hot family and the pixel goal frame are the same imagined world."""
    N, seed = 200, 5
    a = cw.BatchedCraftingWorldEnvOneHot(N, seed=seed)
    b = cw.BatchedCraftingWorldEnv(N, seed=seed)
    a.reset(); b.reset()
    H, W = a.cfg.H, a.cfg.W
    ag = a.goal_agent
    frames = b.render(state=(a.goal_grid[:, :H * W].reshape(N, H, W), ag & 0xFF, (ag >> 8) & 0xFF, (ag >> 16) & 0xFF))
    assert torch.equal(frames, b.desired_goal)


def test_altobs_env_matches_reference_frames_and_oracle(cw):
    """cw_render_alt against frames frozen from the reference AltObs renderer, then the AltObs env end to end."""
    src, frame_t, frames = gu.load_altobs()
    d = gu.load(src)
    B = frames.shape[0]
    env = cw.BatchedCraftingWorldEnvAltObs(B, size=(8, 8), max_steps=d["max_steps"], seed=0, auto_reset=False)
    for i, t in enumerate(frame_t):
        env.load_state(d["grid"][:, t], d["r"][:, t], d["c"][:, t], d["hold"][:, t], d["desired"])
        got = env.observation["observation"].cpu().numpy()
        assert got.dtype == np.int16 and np.array_equal(got, frames[:, i]), t
    assert frames.max() > 255
    # end to end with auto-reset: observation / init / goal frames from the oracle's compact states
    N, seed = 64, 12
    env = cw.BatchedCraftingWorldEnvAltObs(N, size=(6, 6), max_steps=8, seed=seed)
    ob = native.OracleBatch(native.make_config(H=6, W=6, max_steps=8), N, seed=seed)
    pc = compact.Config(H=6, W=6, max_steps=8)
    obs = env.reset(); ob.reset()

    def frames_of(grid, agent):
        return np.stack([compact.render_alt(grid[n, :36].reshape(6, 6), int(agent[n] & 0xFF), int((agent[n] >> 8) & 0xFF),
                                            int((agent[n] >> 16) & 0xFF)) for n in range(N)])
    rng = np.random.RandomState(0)
    for k in range(30):
        a = rng.randint(0, 6, N).astype(np.uint8)
        obs, reward, done, _ = env.step(a)
        o_reward, o_done = ob.step_full(a, auto_reset=True)
        assert np.array_equal(reward.cpu().numpy(), o_reward)
        assert np.array_equal(obs["observation"].cpu().numpy(), frames_of(ob.grid, ob.agent)), k
    goal = []
    for n in range(N):
        _, (g, r, c, h) = compact.reset_env(seed, n, int(ob.episode[n]) - 1, pc, with_goal=True)
        goal.append(compact.render_alt(g, r, c, h))
    assert np.array_equal(obs["desired_goal"].cpu().numpy(), np.stack(goal))
    stacked = cw.BatchedCraftingWorldEnvAltObs(4, size=(6, 6), seed=1, stacked_obs=True)
    assert stacked.reset().shape == (4, 4, 21, 18, 3)


@pytest.mark.parametrize("name", ["variants_onehot_6x6.npz", "variants_flat_8x8.npz"])
def test_variant_envs_match_reference_runs(cw, name):
    """The batched OneHot / Flat mirrors against runs frozen from the reference's own CraftingWorldEnvOneHot and
    CraftingWorldEnvFlat classes (tests/golden/make_golden_variants.py): observation, reward, done at every step."""
    import os
    d = dict(np.load(os.path.join(gu.GOLDEN_DIR, name)))
    size, max_steps, flat, B = int(d["size"]), int(d["max_steps"]), bool(d["flat"]), 3
    cls = cw.BatchedCraftingWorldEnvFlat if flat else cw.BatchedCraftingWorldEnvOneHot
    env = cls(B, size=(size, size), max_steps=max_steps, seed=0, auto_reset=False)
    rep = lambda x: np.repeat(np.asarray(x)[None], B, axis=0)     # the same world B times
    env.load_state(rep(d["grid0"]), rep(d["r0"]), rep(d["c0"]), rep(d["hold0"]), rep(d["desired"]))
    first = env.obs if flat else env.observation["observation"]
    assert np.array_equal(first.cpu().numpy(), rep(d["obs0"]))
    for t, a in enumerate(d["actions"]):
        obs, reward, done, _ = env.step(np.full(B, a, np.uint8))
        got = obs if flat else obs["observation"]
        assert np.array_equal(got.cpu().numpy(), rep(d["obs"][t])), t
        assert reward.cpu().tolist() == [int(d["reward"][t])] * B and done.cpu().tolist() == [bool(d["done"][t])] * B, t


def _random_compact_state(N, size, seed):
    """dense random worlds (p=0.5 per cell, uniform object type), random agent cell and held item"""
    rng = np.random.RandomState(seed)
    stride = (size * size + 15) // 16 * 16
    grid = np.zeros((N, stride), np.uint8)
    cells = rng.randint(0, 9, (N, size * size)) * (rng.rand(N, size * size) < 0.5)
    grid[:, :size * size] = cells
    r, c, h = rng.randint(0, size, N), rng.randint(0, size, N), rng.randint(0, 4, N)
    h[: max(1, N // 3)] = grid[np.arange(N), r * size + c][: max(1, N // 3)].clip(0, 3)   # held item == object under the agent: multiplicity 2
    agent = (r | (c << 8) | (h << 16)).astype(np.int32)
    return grid, agent, r, c, h


@pytest.mark.parametrize("N,size", [(1, 5), (3, 7), (37, 8), (100, 21), (9, 32), (5, 40), (3, 64), (4099, 21)])
def test_onehot_and_altobs_kernels_match_oracle_on_dense_states(cw, N, size):
    """cw_onehot / cw_render_alt (staged, aligned copy-out; work items of several worlds or bands of one world) on dense random
    states of every regime: tiny worlds packed per item, 21x21, worlds split into bands (40x40, 64x64), ragged tails."""
    grid, agent, r, c, h = _random_compact_state(N, size, 1000 + N + size)
    env = cw.BatchedCraftingWorldEnvAltObs(N, size=(size, size), seed=0)
    g, a = torch.from_numpy(grid).cuda(), torch.from_numpy(agent).cuda()
    oh = env.onehot(grid=g, agent=a).cpu().numpy()
    alt = env.render_alt(g, a).cpu().numpy()
    check = range(N) if N <= 128 else list(range(0, N, 97)) + [N - 1]
    for n in check:
        g2 = grid[n, :size * size].reshape(size, size)
        assert np.array_equal(oh[n], ref_shim.compact_to_onehot(g2, int(r[n]), int(c[n]), int(h[n])).astype(np.uint8)), n
        assert np.array_equal(alt[n], compact.render_alt(g2, int(r[n]), int(c[n]), int(h[n]))), n
    # an output carved at every 2- / 4-byte phase of a 16-byte line: the head / tail paths of the copy-out
    for phase in (4, 8, 12):
        raw = torch.full((oh.size + 64,), 0xA5, dtype=torch.uint8, device="cuda")
        view = raw[phase:phase + oh.size]
        env._lib.cw_onehot(__import__("ctypes").byref(env.cfg), g.data_ptr(), a.data_ptr(), view.data_ptr(), N, torch.cuda.current_stream().cuda_stream)
        assert np.array_equal(view.cpu().numpy().reshape(oh.shape), oh), phase
        assert bool((raw[:phase] == 0xA5).all()) and bool((raw[phase + oh.size:] == 0xA5).all()), phase
    for phase in (2, 6, 10, 14):
        nb = alt.size * 2
        raw = torch.full((nb + 64,), 0xA5, dtype=torch.uint8, device="cuda")
        view = raw[phase:phase + nb]
        env._lib.cw_render_alt(__import__("ctypes").byref(env.cfg), g.data_ptr(), a.data_ptr(), view.data_ptr(), N, torch.cuda.current_stream().cuda_stream)
        assert np.array_equal(view.cpu().numpy().view(np.int16).reshape(alt.shape), alt), phase
        assert bool((raw[:phase] == 0xA5).all()) and bool((raw[phase + nb:] == 0xA5).all()), phase
    # unaligned one-hot output: the byte-wise fallback
    raw = torch.full((oh.size + 64,), 0xA5, dtype=torch.uint8, device="cuda")
    view = raw[3:3 + oh.size]
    env._lib.cw_onehot(__import__("ctypes").byref(env.cfg), g.data_ptr(), a.data_ptr(), view.data_ptr(), N, torch.cuda.current_stream().cuda_stream)
    assert np.array_equal(view.cpu().numpy().reshape(oh.shape), oh)


def test_vector_env_facade(cw):
    """gymnasium.vector-shaped facade: call shapes, terminated / truncated split, and every value against the oracle."""
    from oracle import native
    N, size, max_steps, seed = 48, 5, 6, 3
    venv = cw.CraftingWorldVectorEnv(N, size=(size, size), max_steps=max_steps, seed=seed)
    ob = native.OracleBatch(native.make_config(H=size, W=size, max_steps=max_steps), N, seed=seed)
    obs, info = venv.reset(seed=seed)
    o_goal = ob.reset(with_goal=True)
    o_obs = ob.render()
    assert obs["observation"].shape == (N, 20, 20, 3) and info == {}
    assert np.array_equal(obs["observation"].cpu().numpy(), o_obs) and np.array_equal(obs["desired_goal"].cpu().numpy(), o_goal)
    rng = np.random.RandomState(1)
    seen_trunc = seen_term = False
    for k in range(40):
        a = rng.randint(0, 6, N).astype(np.uint8)
        venv.step_async(torch.from_numpy(a).cuda())
        obs, reward, terminated, truncated, info = venv.step_wait()
        o_reward, o_done = ob.step_full(a, auto_reset=True, obs=o_obs)
        assert np.array_equal(reward.cpu().numpy(), o_reward)
        assert np.array_equal(terminated.cpu().numpy(), o_reward == max_steps)
        assert np.array_equal((terminated | truncated).cpu().numpy(), o_done.astype(bool))
        assert not bool((terminated & truncated).any())
        assert np.array_equal(obs["observation"].cpu().numpy(), o_obs), k
        assert np.array_equal(info["achieved_mask"].cpu().numpy().astype(np.uint32), ob.goal & 0xFFFF)
        assert np.array_equal(info["desired_mask"].cpu().numpy().astype(np.uint32), ob.goal >> 16)
        seen_trunc |= bool(truncated.any()); seen_term |= bool(terminated.any())
    assert seen_trunc and set(info) == {"achieved_mask", "desired_mask"}
    nv = cw.CraftingWorldVectorEnv(8, to_numpy=True, size=(5, 5), seed=0)
    o, _ = nv.reset()
    assert isinstance(o["observation"], np.ndarray)


def test_registered_entry_points_construct_the_mirrors(cw):
    """What gym.make would do with the three registrations: import the entry point, call it with the registered kwargs."""
    import importlib
    made = {}
    cw.register_envs(num_envs=16, register=lambda id, entry_point, kwargs: made.__setitem__(id, (entry_point, kwargs)))
    kinds = {}
    for env_id, (entry, kwargs) in made.items():
        mod, cls = entry.split(":")
        env = getattr(importlib.import_module(mod), cls)(**kwargs)
        obs = env.reset()
        kinds[env_id] = type(env).__name__
        assert env.num_envs == 16 and env.stacking is True
        first = obs if isinstance(obs, torch.Tensor) else obs["observation"]
        assert first.shape[0] == 16
    assert kinds == {"craftingworld-b200-v3": "BatchedCraftingWorldEnv", "craftingworldflat-b200-v3": "BatchedCraftingWorldEnvFlat",
                     "craftingworldonehot-b200-v3": "BatchedCraftingWorldEnvOneHot"}


def test_init_observation_vector_keeps_the_initial_agent_channels(cw):
    """`observation_vector['init_observation']` is INIT_OBS_VECTOR (ray.py:183-187): object AND agent / holding channels as they
    were at reset -- in pixel mode too, through auto-resets, and for the compact step kernel."""
    from oracle import native
    N, size, max_steps, seed = 300, 6, 5, 12
    for mode in ("pixels", "compact"):
        env = cw.BatchedCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, seed=seed, obs_mode=mode)
        ob = native.OracleBatch(native.make_config(H=size, W=size, max_steps=max_steps), N, seed=seed)
        env.reset(); ob.reset()
        init_agent = ob.agent.copy()
        rng = np.random.RandomState(0)
        for k in range(17):
            a = rng.randint(0, 6, N).astype(np.uint8)
            env.step(torch.from_numpy(a).cuda())
            _, o_done = ob.step_full(a, auto_reset=True)
            init_agent[o_done == 1] = ob.agent[o_done == 1]         # a re-seeded world starts a new INIT_OBS_VECTOR
        assert np.array_equal(env.init_agent.cpu().numpy().astype(np.uint32), init_agent), mode
        vec = env.observation_vector["init_observation"].cpu().numpy()
        want = np.zeros((N, size, size, 12), np.uint8)
        ig = ob.init_grid[:, :size * size].reshape(N, size, size)
        for ch in range(8):
            want[..., ch] = ig == ch + 1
        want[np.arange(N), init_agent & 0xFF, (init_agent >> 8) & 0xFF, 8] = 1
        assert np.array_equal(vec, want), mode
        assert (env.agent.cpu().numpy().astype(np.uint32) != init_agent).any()          # (the current agent words differ)


def test_altobs_render_of_foreign_states(cw):
    """AltObs `render(state)` (craftingworld_altobs.py:489): frames of states that are NOT the env's own (a different count
    of them, too), against the frames frozen from the reference's AltObs renderer."""
    src, frame_t, frames = gu.load_altobs()
    d = gu.load(src)
    env = cw.BatchedCraftingWorldEnvAltObs(3, size=(8, 8), seed=0)
    env.reset()
    before = env.render().clone()
    for i, t in enumerate(frame_t[:4]):
        got = env.render((d["grid"][:, t], d["r"][:, t], d["c"][:, t], d["hold"][:, t])).cpu().numpy()
        assert got.shape == frames[:, i].shape and np.array_equal(got, frames[:, i]), t
    assert torch.equal(env.render(), before)                       # the env's own worlds are untouched


def test_out_of_range_actions_of_wide_dtypes_stay_no_ops(cw):
    """260 or -252 must not wrap modulo 256 onto 'pickup': an out-of-range action is a no-op that still advances step_num."""
    N = 64
    a = cw.BatchedCraftingWorldEnv(N, size=(5, 5), max_steps=50, seed=4, auto_reset=False)
    b = cw.BatchedCraftingWorldEnv(N, size=(5, 5), max_steps=50, seed=4, auto_reset=False)
    a.reset(); b.reset()
    for k in range(10):                                             # put some agents onto pickupable objects
        acts = torch.randint(0, 4, (N,), device="cuda")
        a.step(acts); b.step(acts)
    wide = torch.tensor([260, -252, 6, 1000] * (N // 4), device="cuda", dtype=torch.int64)
    a.step(wide)
    b.step(torch.full((N,), 6, device="cuda", dtype=torch.uint8))
    for key in ("grid", "agent", "goal", "t", "reward"):
        assert torch.equal(getattr(a, key), getattr(b, key)), key
    assert bool((a.t == 11).all())


def test_frame_policy_consumer_reads_every_byte(cw):
    """cw_frame_policy (the stand-in device consumer of the closed-loop bench leg): its actions equal the NumPy restatement of
    the hash, and flipping ONE byte of a frame changes the hash."""
    for size, N in ((21, 300), (5, 33), (32, 17)):
        env = cw.BatchedCraftingWorldEnv(N, size=(size, size), seed=size)
        obs = env.reset()["observation"]
        got = env.frame_policy().cpu().numpy()
        words = obs.cpu().numpy().reshape(N, -1).view(np.uint32).astype(np.uint64)
        k = (2 * np.arange(words.shape[1], dtype=np.uint64) + 1)
        h = ((words * k) & 0xFFFFFFFF).sum(axis=1) & 0xFFFFFFFF
        want = ((h ^ (h >> 16)) & 0xFFFF) % 6
        assert np.array_equal(got, want.astype(np.uint8)), size
    obs2 = obs.clone()
    obs2[3, -1, -1, 2] ^= 1                                        # the very last byte of world 3
    w2 = obs2.cpu().numpy().reshape(N, -1).view(np.uint32).astype(np.uint64)
    assert (((w2 * k) & 0xFFFFFFFF).sum(axis=1) & 0xFFFFFFFF)[3] != h[3]
    got2 = env.frame_policy(obs2).cpu().numpy()
    want2 = ((( ((w2 * k) & 0xFFFFFFFF).sum(axis=1) & 0xFFFFFFFF) ^ ((((w2 * k) & 0xFFFFFFFF).sum(axis=1) & 0xFFFFFFFF) >> 16)) & 0xFFFF) % 6
    assert np.array_equal(got2, want2.astype(np.uint8))


def test_integration_md_ctypes_stub_runs(cw):
    """The raw ctypes binding printed in INTEGRATION.md section 4 is executed verbatim and checked against the facade."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    section = text[text.index("## 4. The raw ctypes stub"):text.index("## 5. Build")]
    code = re.search(r"```python\n(.*?)```", section, flags=re.S).group(1)
    cwd = os.getcwd()
    os.chdir(root)
    try:
        ns = {}
        exec(compile(code, "INTEGRATION.md", "exec"), ns)
    finally:
        os.chdir(cwd)
    torch.cuda.synchronize()
    assert ns["rc"] == 0
    # same seed / ids through the facade: identical worlds, frames and step results
    env = cw.BatchedCraftingWorldEnv(ns["N"], seed=1234, goal_images=False)
    env.reset()
    _, reward, done, _ = env.step(ns["actions"])
    assert torch.equal(env.grid, ns["grid"]) and torch.equal(env.agent, ns["agent"]) and torch.equal(env.obs, ns["obs"])
    assert torch.equal(reward, ns["reward"]) and torch.equal(done, ns["done"].bool())
