"""GPU (-m gpu): the CUDA path, called through the Python facade and hence the C ABI, against
(i) the frozen reference traces in tests/golden and (ii) the C oracle on the same seeded inputs.  Bit-exact."""
import numpy as np
import pytest
import torch

from oracle import native, ref_shim
from tests import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cw():
    import gym_craftingworld_b200 as pkg
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return pkg


def make_env(cw, d, **kw):
    B = d["actions"].shape[0]
    kw.setdefault("auto_reset", False)
    kw.setdefault("goal_images", False)
    env = cw.BatchedCraftingWorldEnv(B, size=(d["W"], d["H"]), max_steps=d["max_steps"],
                                     reward_style=("subset" if d["subset"] else None), seed=0, **kw)
    env.load_state(d["grid0"], d["r0"], d["c0"], d["hold0"], d["desired"])
    return env


def assert_state_equals_golden(env, d, t, where):
    s = env.export_state()
    assert np.array_equal(s["grid"], d["grid"][:, t]), where
    assert np.array_equal(s["r"], d["r"][:, t]) and np.array_equal(s["c"], d["c"][:, t]), where
    assert np.array_equal(s["hold"], d["hold"][:, t]), where
    assert np.array_equal(s["achieved"], d["achieved"][:, t]), where


@pytest.mark.parametrize("name", gu.golden_files())
def test_fused_step_render_matches_reference_trace(cw, name):
    """cw_step_render (one fused launch per step) reproduces the reference trace: state, reward, done, pixels."""
    d = gu.load(name)
    env = make_env(cw, d)
    B, T = d["actions"].shape
    fidx = {int(t): i for i, t in enumerate(d["frame_t"])}
    assert np.array_equal(env.obs.cpu().numpy(), d["frame0"])
    for t in range(T):
        obs, reward, done, info = env.step(torch.from_numpy(d["actions"][:, t]).cuda())
        where = f"{name} step {t}"
        assert np.array_equal(reward.cpu().numpy(), d["reward"][:, t]), where
        assert np.array_equal(done.cpu().numpy(), d["done"][:, t].astype(bool)), where
        assert_state_equals_golden(env, d, t, where)
        frames = obs["observation"].cpu().numpy()
        assert [gu.crc(f) for f in frames] == list(d["frame_crc"][:, t]), where
        if t in fidx:
            assert np.array_equal(frames, d["frames"][:, fidx[t]]), where
    assert np.array_equal(info["achieved_goal"].cpu().numpy(),
                          (d["achieved"][:, -1, None] >> np.arange(9)) & 1)


@pytest.mark.parametrize("name", ["quirks_5x5.npz", "quirks_5x5_subset.npz", "dense_8x8.npz", "dense_21x21.npz", "dense_32x32.npz"])
def test_compact_step_then_render_matches_reference_trace(cw, name):
    """cw_step (thread-per-world kernel) + cw_render as separate launches give the same trace."""
    d = gu.load(name)
    env = make_env(cw, d, obs_mode="compact")
    B, T = d["actions"].shape
    for t in range(T):
        obs, reward, done, _ = env.step(d["actions"][:, t])
        where = f"{name} step {t}"
        assert np.array_equal(reward.cpu().numpy(), d["reward"][:, t]), where
        assert np.array_equal(done.cpu().numpy(), d["done"][:, t].astype(bool)), where
        assert_state_equals_golden(env, d, t, where)
        assert np.array_equal(obs["observation"].cpu().numpy(), d["grid"][:, t]), where
        if t % 16 == 0 or t == T - 1:
            frames = env.render().cpu().numpy()
            assert [gu.crc(f) for f in frames] == list(d["frame_crc"][:, t]), where


@pytest.mark.parametrize("name", ["dense_5x5_subset.npz", "dense_21x21.npz", "sampled_21x21.npz"])
def test_rollout_matches_reference_trace(cw, name):
    """cw_rollout: the whole T-step action tape in ONE launch reproduces every per-step reward/done and the final state."""
    d = gu.load(name)
    env = make_env(cw, d, obs_mode="compact")
    T = d["actions"].shape[1]
    rew, dn = env.rollout(np.ascontiguousarray(d["actions"].T))
    assert np.array_equal(rew.cpu().numpy(), d["reward"].T)
    assert np.array_equal(dn.cpu().numpy(), d["done"].T.astype(bool))
    assert_state_equals_golden(env, d, T - 1, name)
    assert np.array_equal(env.t.cpu().numpy(), np.full(d["actions"].shape[0], T))


def oracle_for(env, seed):
    cfg = native.make_config(H=env.cfg.H, W=env.cfg.W, max_steps=env.cfg.max_steps, subset_reward=bool(env.cfg.subset_reward),
                             stacking=bool(env.cfg.stacking), selected=tuple(env.cfg.selected[i] for i in range(env.cfg.n_selected)),
                             number_of_tasks=env.cfg.number_of_tasks)
    return native.OracleBatch(cfg, env.num_envs, seed=seed, env_id_base=env.env_id_base)


def assert_env_equals_oracle(env, ob, where=""):
    assert np.array_equal(env.grid.cpu().numpy(), ob.grid), where + " grid"
    assert np.array_equal(env.init_grid.cpu().numpy(), ob.init_grid), where + " init_grid"
    assert np.array_equal(env.agent.cpu().numpy().astype(np.uint32), ob.agent), where + " agent"
    assert np.array_equal(env.goal.cpu().numpy().astype(np.uint32), ob.goal), where + " goal"
    assert np.array_equal(env.t.cpu().numpy(), ob.t), where + " t"
    assert np.array_equal(env.episode.cpu().numpy().astype(np.uint32), ob.episode), where + " episode"


@pytest.mark.parametrize("size,kw", [
    (21, {}),
    (32, dict(selected_tasks=['ChopTree', 'BuildHouse'], number_of_tasks=2)),
    (8, dict(stacking=False, selected_tasks=['MakeBread', 'EatBread', 'GoToHouse', 'MoveSticks'])),
    (4, dict(number_of_tasks=3)),
    (64, {}),
])
def test_reset_bit_exact_vs_oracle(cw, size, kw):
    """Philox reset: placement, tasks, first frame, imagine_obs goal frame and INIT_OBS copy, over 3 episodes."""
    N, seed = 300, 20240607
    env = cw.BatchedCraftingWorldEnv(N, size=(size, size), seed=seed, env_id_base=5000, **kw)
    ob = oracle_for(env, seed)
    for ep in range(3):
        obs = env.reset()
        goal = ob.reset(with_goal=True)
        assert_env_equals_oracle(env, ob, f"{size}x{size} episode {ep}")
        assert np.array_equal(obs["observation"].cpu().numpy(), ob.render())
        assert np.array_equal(obs["desired_goal"].cpu().numpy(), goal)
        assert np.array_equal(obs["init_observation"].cpu().numpy(), ob.render())
        assert obs["achieved_goal"] is obs["observation"]                       # aliasing as upstream (ray.py:194-196)


def test_masked_reset(cw):
    N, seed = 257, 9
    env = cw.BatchedCraftingWorldEnv(N, seed=seed)
    ob = oracle_for(env, seed)
    env.reset(); ob.reset(with_goal=True)
    mask = (np.arange(N) % 3 == 0).astype(np.uint8)
    before = env.obs.clone()
    env.reset(mask=torch.from_numpy(mask).cuda())
    ob.reset(mask=mask)
    assert_env_equals_oracle(env, ob, "masked reset")
    after = env.obs.cpu().numpy()
    assert np.array_equal(after, ob.render())
    assert np.array_equal(after[mask == 0], before.cpu().numpy()[mask == 0])


@pytest.mark.parametrize("size,N,max_steps,K", [(21, 1000, 25, 120), (5, 513, 7, 60), (32, 200, 30, 70)])
def test_autoreset_trajectory_vs_oracle(cw, size, N, max_steps, K):
    """Fused step + auto-reset + render (+ goal / init frames + stats) against the C oracle's cwo_step_full,
    every step: reward, done; every 10th step and at the end: full state and all three frame buffers."""
    seed = 77
    env = cw.BatchedCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, seed=seed, auto_reset=True)
    ob = oracle_for(env, seed)
    env.reset()
    o_goal = ob.reset(with_goal=True)
    o_obs = ob.render()
    o_init = o_obs.copy()
    rng = np.random.RandomState(5)
    for k in range(K):
        a = rng.randint(0, 6, N).astype(np.uint8)
        obs, reward, done, _ = env.step(torch.from_numpy(a).cuda())
        new_goal = np.zeros_like(o_goal)
        o_reward, o_done = ob.step_full(a, auto_reset=True, obs=o_obs, goal_obs=new_goal)
        o_goal[o_done == 1] = new_goal[o_done == 1]
        o_init[o_done == 1] = o_obs[o_done == 1]
        assert np.array_equal(reward.cpu().numpy(), o_reward), f"step {k}"
        assert np.array_equal(done.cpu().numpy(), o_done.astype(bool)), f"step {k}"
        if k % 10 == 0 or k == K - 1:
            assert_env_equals_oracle(env, ob, f"step {k}")
            assert np.array_equal(obs["observation"].cpu().numpy(), o_obs), f"step {k} obs"
            assert np.array_equal(obs["desired_goal"].cpu().numpy(), o_goal), f"step {k} goal frame"
            assert np.array_equal(obs["init_observation"].cpu().numpy(), o_init), f"step {k} init frame"
    assert np.array_equal(env.stats.cpu().numpy(), ob.stats)
    st = env.episode_stats()
    assert st["episodes"] == int(ob.stats[0]) > 0


def test_autoreset_compact_and_rollout_vs_oracle(cw):
    """Thread-per-world kernel with warp-cooperative auto-reset: single steps and a K-step rollout launch."""
    N, seed, K = 2000, 31, 150
    acts = np.random.RandomState(2).randint(0, 6, (K, N)).astype(np.uint8)
    ob = None
    for mode in ("steps", "rollout"):
        env = cw.BatchedCraftingWorldEnv(N, size=(8, 8), max_steps=12, seed=seed, obs_mode="compact")
        ob = oracle_for(env, seed)
        env.reset(); ob.reset()
        o_rew = np.zeros((K, N), np.int32); o_dn = np.zeros((K, N), np.uint8)
        for k in range(K):
            o_rew[k], o_dn[k] = ob.step_full(acts[k], auto_reset=True)
        if mode == "steps":
            for k in range(K):
                _, reward, done, _ = env.step(acts[k])
                assert np.array_equal(reward.cpu().numpy(), o_rew[k]) and np.array_equal(done.cpu().numpy(), o_dn[k].astype(bool))
        else:
            rew, dn = env.rollout(acts)
            assert np.array_equal(rew.cpu().numpy(), o_rew) and np.array_equal(dn.cpu().numpy(), o_dn.astype(bool))
        assert_env_equals_oracle(env, ob, mode)
        assert np.array_equal(env.stats.cpu().numpy(), ob.stats), mode


def test_fixed_init_state_pool(cw):
    """fixed_init_state=n (ray.py:116-118, 630-644): every reset copies one of n pre-sampled worlds."""
    N, seed, n_fixed = 400, 123, 5
    env = cw.BatchedCraftingWorldEnv(N, size=(9, 9), fixed_init_state=n_fixed, seed=seed)
    pool_g = env._fixed_grid.cpu().numpy()
    pool_a = env._fixed_agent.cpu().numpy().astype(np.uint32)
    # the pool itself is a Philox sample_state of streams FIXED_POOL_ID_BASE + i
    pool_oracle = native.OracleBatch(native.make_config(H=9, W=9), n_fixed, seed=seed, env_id_base=1 << 62)
    pool_oracle.reset()
    assert np.array_equal(pool_g, pool_oracle.grid) and np.array_equal(pool_a, pool_oracle.agent)
    ob = oracle_for(env, seed)
    native.set_fixed_pool(pool_g, pool_a)
    try:
        for ep in range(2):
            env.reset(); goal = ob.reset(with_goal=True)
            assert_env_equals_oracle(env, ob, f"fixed pool episode {ep}")
            assert np.array_equal(env.desired_goal.cpu().numpy(), goal)
        grids = env.grid.cpu().numpy()
        assert all(any(np.array_equal(g, p) for p in pool_g) for g in grids)
        assert len({g.tobytes() for g in grids}) == n_fixed
    finally:
        native.set_fixed_pool(None)


def test_imagine_on_injected_dense_worlds(cw):
    """cw_imagine on injected (dense, possibly holding) worlds equals the oracle's imagine_obs + render."""
    d = gu.load("dense_8x8.npz")
    env = make_env(cw, d, goal_images=True)
    env2 = oracle_for(env, 0)
    env2.load_state(d["grid0"], d["r0"], d["c0"], d["hold0"], d["desired"])
    ig, ia = env2.imagine()
    tmp = native.OracleBatch(env2.cfg, env2.N)
    tmp.grid[:], tmp.agent[:] = ig, ia
    assert np.array_equal(env.desired_goal.cpu().numpy(), tmp.render())
    assert np.array_equal(env.init_obs.cpu().numpy(), d["frame0"])


def test_onehot_observation_vector(cw):
    d = gu.load("dense_8x8.npz")
    env = make_env(cw, d)
    for t in range(8):
        env.step(d["actions"][:, t])
    oh = env.onehot().cpu().numpy()
    s = env.export_state()
    for b in range(env.num_envs):
        want = ref_shim.compact_to_onehot(s["grid"][b], int(s["r"][b]), int(s["c"][b]), int(s["hold"][b]))
        assert np.array_equal(oh[b], want.astype(np.uint8)), b
    ov = env.observation_vector
    assert ov["observation"].shape == (env.num_envs, 8, 8, 12) and ov["desired_goal"].shape == (env.num_envs, 9)


def test_render_arbitrary_states(cw):
    d = gu.load("dense_21x21.npz")
    env = cw.BatchedCraftingWorldEnv(4, seed=0)
    t = 31
    out = env.render(state=(d["grid"][:, t], d["r"][:, t], d["c"][:, t], d["hold"][:, t])).cpu().numpy()
    assert [gu.crc(f) for f in out] == list(d["frame_crc"][:, t])


def test_sharding_does_not_change_worlds(cw):
    """Worlds are keyed by GLOBAL id: two half-size envs with env_id_base offsets == one full env."""
    N, seed, K = 512, 4242, 40
    acts = np.random.RandomState(0).randint(0, 6, (K, N)).astype(np.uint8)
    full = cw.BatchedCraftingWorldEnv(N, size=(6, 6), max_steps=9, seed=seed)
    a = cw.BatchedCraftingWorldEnv(200, size=(6, 6), max_steps=9, seed=seed, env_id_base=0)
    b = cw.BatchedCraftingWorldEnv(312, size=(6, 6), max_steps=9, seed=seed, env_id_base=200)
    for e in (full, a, b):
        e.reset()
    for k in range(K):
        full.step(acts[k]); a.step(acts[k, :200]); b.step(acts[k, 200:])
    for key in ("grid", "agent", "goal", "t", "episode", "obs", "desired_goal", "init_obs"):
        whole = getattr(full, key).cpu().numpy()
        parts = np.concatenate([getattr(a, key).cpu().numpy(), getattr(b, key).cpu().numpy()])
        assert np.array_equal(whole, parts), key
    assert np.array_equal(full.stats.cpu().numpy(), (a.stats + b.stats).cpu().numpy())


def test_invalid_actions(cw):
    env = cw.BatchedCraftingWorldEnv(8, size=(5, 5), seed=1, auto_reset=False)
    env.reset()
    before = env.export_state()
    _, reward, done, _ = env.step(torch.full((8,), 9, dtype=torch.uint8, device="cuda"))
    after = env.export_state()
    assert (reward.cpu().numpy() == -1).all() and not done.any()
    assert np.array_equal(before["grid"], after["grid"]) and (after["t"] == 1).all()
    strict = cw.BatchedCraftingWorldEnv(8, size=(5, 5), seed=1, validate_actions=True)
    strict.reset()
    with pytest.raises(IndexError):
        strict.step(np.full(8, 6))
    with pytest.raises(ValueError):
        env.step(np.zeros(3, np.uint8))


def test_api_surface_and_reward_helpers(cw):
    env = cw.BatchedCraftingWorldEnv(16, seed=3)
    assert env.action_space.n == 6 and env.observation_space["observation"].shape == (84, 84, 3)
    assert env.observation_vector_space["observation"].shape == (21, 21, 12)
    obs = env.reset()
    assert set(obs) == {"observation", "desired_goal", "achieved_goal", "init_observation"}
    assert obs["observation"].shape == (16, 84, 84, 3) and obs["observation"].dtype == torch.uint8
    obs2, reward, done, info = env.step(torch.zeros(16, dtype=torch.int64, device="cuda"))
    assert obs2["observation"] is obs["observation"]                               # owned + mutated in place
    assert set(info) == {"task_success", "desired_goal", "achieved_goal"} and info["desired_goal"].shape == (16, 9)
    assert reward.dtype == torch.int32 and done.dtype == torch.bool
    assert env.seed(5) == [5]
    ach = torch.tensor([[0, 1, 0, 0, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0, 0, 0, 0]], device="cuda")
    des = torch.tensor([[0, 1, 0, 0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0, 0, 0, 0]], device="cuda")
    assert env.compute_reward_equal(ach, des).tolist() == [300, -1]
    assert env.compute_reward_subset(ach, des).tolist() == [300, 300]
    assert env.compute_reward(ach[0], des[0], None).tolist() == [300]


def test_step_is_cuda_graph_capturable(cw):
    N, K = 256, 16
    acts = torch.from_numpy(np.random.RandomState(3).randint(0, 6, (K, N)).astype(np.uint8)).cuda()
    ref = cw.BatchedCraftingWorldEnv(N, size=(7, 7), max_steps=10, seed=8)
    env = cw.BatchedCraftingWorldEnv(N, size=(7, 7), max_steps=10, seed=8)
    ref.reset(); env.reset()
    for k in range(K):
        ref.step(acts[k])
    s = torch.cuda.Stream()                                   # (ref's steps above already warmed the launch path)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for k in range(K):
                env.step(acts[k])
    g.replay()
    torch.cuda.synchronize()
    for key in ("grid", "agent", "goal", "t", "episode", "obs", "desired_goal"):
        assert torch.equal(getattr(env, key), getattr(ref, key)), key


@pytest.mark.parametrize("name", gu.golden_files())
def test_incremental_render_matches_reference_trace(cw, name):
    """cw_step_render_edit -- the reference's own render_edit (ray.py:522-557) on the device frame: only the <= 2 changed cells
    are rewritten each step -- reproduces the reference trace: state, reward, done and every pixel of every step."""
    d = gu.load(name)
    env = make_env(cw, d, render="incremental")
    B, T = d["actions"].shape
    fidx = {int(t): i for i, t in enumerate(d["frame_t"])}
    for t in range(T):
        obs, reward, done, _ = env.step(torch.from_numpy(d["actions"][:, t]).cuda())
        where = f"{name} step {t}"
        assert np.array_equal(reward.cpu().numpy(), d["reward"][:, t]) and np.array_equal(done.cpu().numpy(), d["done"][:, t].astype(bool)), where
        assert_state_equals_golden(env, d, t, where)
        frames = obs["observation"].cpu().numpy()
        assert [gu.crc(f) for f in frames] == list(d["frame_crc"][:, t]), where
        if t in fidx:
            assert np.array_equal(frames, d["frames"][:, fidx[t]]), where


@pytest.mark.parametrize("N,size,max_steps", [(700, 7, 5), (4096, 21, 40), (300, 32, 12)])
def test_incremental_render_with_auto_reset_matches_full_render(cw, N, size, max_steps):
    """Incremental and full rendering side by side with auto-reset: identical state, reward, done, statistics and identical
    observation / goal / init frames after every step (the finished worlds are re-seeded by the masked reset launch)."""
    K = 3 * max_steps + 7
    acts = torch.from_numpy(np.random.RandomState(N).randint(0, 6, (K, N)).astype(np.uint8)).cuda()
    full = cw.BatchedCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, seed=17)
    inc = cw.BatchedCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, seed=17, render="incremental")
    o1, o2 = full.reset(), inc.reset()
    for key in ("observation", "desired_goal", "init_observation"):
        assert torch.equal(o1[key], o2[key]), key
    for k in range(K):
        o1, r1, d1, _ = full.step(acts[k])
        o2, r2, d2, _ = inc.step(acts[k])
        assert torch.equal(r1, r2) and torch.equal(d1, d2), k
        for key in ("observation", "desired_goal", "init_observation"):
            assert torch.equal(o1[key], o2[key]), (key, k)
    for key in ("grid", "init_grid", "agent", "goal", "t", "episode", "stats"):   # (stats: the 16 replicas are filled differently)
        assert torch.equal(getattr(full, key), getattr(inc, key)), key
    assert int(full.stats[0]) >= N


@pytest.mark.parametrize("N,size,max_steps,ring,K,graph", [
    (64, 7, 3, 1, 40, True),          # tiny launches: several chain positions co-resident, one frame buffer, resets every <= 3 steps
    (300, 7, 2, 3, 48, True),         # a world can be re-seeded in consecutive positions (goal / init frames have no ring)
    (4096, 21, 20, 4, 128, True),     # the bench shape: one CTA wave per launch
    (5000, 21, 300, 2, 64, False),    # chained launches outside a graph
    (40000, 21, 10, 2, 16, True),     # persistent launches (more groups than CTAs)
    (700, 32, 6, 2, 24, True),        # multi-chunk frames
])
def test_chained_steps_match_unchained(cw, N, size, max_steps, ring, K, graph):
    """cw_step_render_chained (per-group dataflow between consecutive launches) == the same steps launched one by one."""
    acts = torch.from_numpy(np.random.RandomState(N + K).randint(0, 6, (K, N)).astype(np.uint8)).cuda()
    kw = dict(size=(size, size), max_steps=max_steps, seed=21, obs_buffers=ring)
    ref = cw.BatchedCraftingWorldEnv(N, **kw)
    env = cw.BatchedCraftingWorldEnv(N, **kw)
    ref.reset(); env.reset()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        env.step(acts[0]); env.step(acts[0], chain_pos=0)       # warm both launch paths outside the capture
        ref.step(acts[0]); ref.step(acts[0])
        s.synchronize()
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for k in range(K):
                    env.step(acts[k], chain_pos=k)
        for rep in range(3):                                      # every replay re-opens the chain at position 0
            if graph:
                g.replay()
            else:
                for k in range(K):
                    env.step(acts[k], chain_pos=k)
            for k in range(K):
                ref.step(acts[k])
            s.synchronize()
            for key in ("grid", "init_grid", "agent", "goal", "t", "episode", "reward", "done", "desired_goal", "init_obs", "stats_raw"):
                assert torch.equal(getattr(env, key), getattr(ref, key)), (key, rep)
            for b in range(ring):
                assert torch.equal(env._obs_ring[b], ref._obs_ring[b]), ("frame buffer", b, rep)


@pytest.mark.parametrize("N,size,max_steps,K,graph", [
    (96, 5, 2, 40, True),             # tiny launches, a reset every <= 2 steps: many positions co-resident
    (65536, 21, 30, 128, True),       # the BASELINE config-3 shape
    (5000, 21, 300, 64, False),       # chained launches outside a graph, worlds not a multiple of 32 / 128
    (33, 8, 1, 20, True),             # every world re-seeded in every position
])
def test_chained_compact_steps_match_unchained(cw, N, size, max_steps, K, graph):
    """cw_step_chained (consecutive launches linked per warp of 32 worlds) == the same steps as ordinary cw_step launches,
    per-step reward / done included (each position writes its own row)."""
    acts = torch.from_numpy(np.random.RandomState(N + K).randint(0, 6, (K, N)).astype(np.uint8)).cuda()
    kw = dict(size=(size, size), max_steps=max_steps, seed=23, obs_mode="compact")
    ref = cw.BatchedCraftingWorldEnv(N, **kw)
    env = cw.BatchedCraftingWorldEnv(N, **kw)
    ref.reset(); env.reset()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        env.step(acts[0]); env.step(acts[0], chain_pos=0)
        ref.step(acts[0]); ref.step(acts[0])
        s.synchronize()
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for k in range(K):
                    env.step(acts[k], chain_pos=k)
        for rep in range(3):
            if graph:
                g.replay()
            else:
                for k in range(K):
                    env.step(acts[k], chain_pos=k)
            for k in range(K):
                ref.step(acts[k])
            s.synchronize()
            for key in ("grid", "init_grid", "agent", "goal", "t", "episode", "reward", "done", "stats_raw"):
                assert torch.equal(getattr(env, key), getattr(ref, key)), (key, rep)
        obs = env.step(acts[0])[0]; ref.step(acts[0])             # an ordinary step after the chain
        assert isinstance(obs, cw.env.CompactObs)                 # lazy views: step() itself must launch nothing but the step kernel
        assert torch.equal(obs["desired_goal"], (ref.goal >> 16) & 0xFFFF) and torch.equal(obs["achieved_goal"], ref.goal & 0xFFFF)
        s.synchronize()
        for key in ("grid", "agent", "goal", "t", "episode", "reward", "done", "stats_raw"):
            assert torch.equal(getattr(env, key), getattr(ref, key)), key


def test_chained_step_rejects_bad_arguments(cw):
    env = cw.BatchedCraftingWorldEnv(8, size=(5, 5), seed=1, obs_mode="onehot")
    env.reset()
    with pytest.raises(ValueError):
        env.step(torch.zeros(8, dtype=torch.uint8, device="cuda"), chain_pos=0)
    env = cw.BatchedCraftingWorldEnv(8, size=(5, 5), seed=1)
    env.reset()
    with pytest.raises(ValueError):
        env.step(torch.zeros(8, dtype=torch.uint8, device="cuda"), chain_pos=1024)


def test_host_env_matches_oracle(cw):
    """The host-buffer API (cw_host_*): NumPy in / NumPy out, sliced + pipelined inside the library."""
    N, seed, K = 3000, 55, 30
    env = cw.HostCraftingWorldEnv(N, size=(21, 21), max_steps=15, seed=seed)
    cfg = native.make_config(H=21, W=21, max_steps=15)
    ob = native.OracleBatch(cfg, N, seed=seed)
    obs = env.reset()
    goal = ob.reset(with_goal=True)
    o_obs = ob.render()
    assert np.array_equal(obs["observation"], o_obs) and np.array_equal(obs["desired_goal"], goal)
    rng = np.random.RandomState(1)
    for k in range(K):
        a = rng.randint(0, 6, N)
        obs, reward, done, _ = env.step(a)
        o_reward, o_done = ob.step_full(a.astype(np.uint8), auto_reset=True, obs=o_obs)
        assert np.array_equal(reward, o_reward) and np.array_equal(done, o_done.astype(bool)), k
        assert np.array_equal(obs["observation"], o_obs), k
    assert np.array_equal(env.stats(), ob.stats)
    env.close()


@pytest.mark.parametrize("mode", ["full", "chained", "incremental"])
def test_no_out_of_bounds_writes_guard_bands(cw, mode):
    """compute-sanitizer is closed on this pool, so out-of-bounds writes are hunted with canaries: every buffer the
    kernels write is carved out of a larger allocation whose guard bands must stay untouched (odd sizes, auto-reset,
    goal + init frames, multi-chunk frames at 64x64, tail groups) -- for independent launches, chained launches and the
    incremental (render_edit + work-list re-seed) path."""
    GUARD = 4096
    for size, N in ((21, 1031), (5, 77), (64, 37), (32, 300)):
        env = cw.BatchedCraftingWorldEnv(N, size=(size, size), max_steps=9, seed=size,
                                         render="incremental" if mode == "incremental" else "full")
        if mode == "incremental":
            env._edit_scratch = torch.zeros(N + 2, dtype=torch.int32, device="cuda")
        if mode == "chained":
            from gym_craftingworld_b200 import _lib as _l
            env._chain = torch.zeros(_l.CHAIN_MAX_POS + N, dtype=torch.int32, device="cuda")
        pads = {}
        names = ["grid", "init_grid", "agent", "goal", "t", "episode", "reward", "_done_u8", "obs", "desired_goal", "init_obs"]
        names += ["_edit_scratch"] if mode == "incremental" else (["_chain"] if mode == "chained" else [])
        for name in names:
            old = getattr(env, name)
            nbytes = old.numel() * old.element_size()
            raw = torch.full((nbytes + 2 * GUARD,), 0xA5, dtype=torch.uint8, device="cuda")
            view = raw[GUARD:GUARD + nbytes].view(old.dtype).view(old.shape)
            view.zero_()
            setattr(env, name, view)
            pads[name] = raw
        env.done = env._done_u8.view(torch.bool)
        env._obs_ring = [env.obs]
        env._refresh_state_struct()
        env.reset()
        acts = torch.randint(0, 6, (40, N), device="cuda", dtype=torch.uint8)
        for k in range(40):
            env.step(acts[k], chain_pos=k if mode == "chained" else None)
        env.render()
        torch.cuda.synchronize()
        for name, raw in pads.items():
            assert bool((raw[:GUARD] == 0xA5).all()) and bool((raw[-GUARD:] == 0xA5).all()), f"{name} guard band hit at {size}x{size}"
        ob = oracle_for(env, size)                     # and the run itself is still correct
        ob.reset()
        o_obs = ob.render()
        for k in range(40):
            ob.step_full(acts[k].cpu().numpy(), auto_reset=True, obs=o_obs)
        assert_env_equals_oracle(env, ob, f"guarded {size}")
        assert np.array_equal(env.obs.cpu().numpy(), o_obs)


def test_no_out_of_bounds_writes_guard_bands_chained_compact(cw):
    """The same hunt for cw_step_chained (one-warp CTAs, per-warp marks, pre-drawn reset records written after the mark): odd
    world counts, every buffer the kernel writes between guard bands."""
    from gym_craftingworld_b200 import _lib as _l
    GUARD = 4096
    for size, N in ((21, 1031), (5, 77), (32, 300)):
        env = cw.BatchedCraftingWorldEnv(N, size=(size, size), max_steps=7, seed=size, obs_mode="compact")
        env._chain = torch.zeros(_l.CHAIN_MAX_POS + (N + 31) // 32, dtype=torch.int32, device="cuda")     # the documented minimum
        pads = {}
        for name in ["grid", "init_grid", "agent", "goal", "t", "episode", "reward", "_done_u8", "reset_rec", "reset_list", "init_agent", "_chain"]:
            old = getattr(env, name)
            nbytes = old.numel() * old.element_size()
            raw = torch.full((nbytes + 2 * GUARD,), 0xA5, dtype=torch.uint8, device="cuda")
            view = raw[GUARD:GUARD + nbytes].view(old.dtype).view(old.shape)
            view.zero_()
            setattr(env, name, view)
            pads[name] = raw
        env.done = env._done_u8.view(torch.bool)
        env._refresh_state_struct()
        env.reset()
        acts = torch.randint(0, 6, (60, N), device="cuda", dtype=torch.uint8)
        for k in range(60):
            env.step(acts[k], chain_pos=k)
        torch.cuda.synchronize()
        for name, raw in pads.items():
            assert bool((raw[:GUARD] == 0xA5).all()) and bool((raw[-GUARD:] == 0xA5).all()), f"{name} guard band hit at {size}x{size}"
        ob = oracle_for(env, size)
        ob.reset()
        for k in range(60):
            ob.step_full(acts[k].cpu().numpy(), auto_reset=True)
        assert_env_equals_oracle(env, ob, f"guarded chained compact {size}")


@pytest.mark.parametrize("size,N,max_steps", [(21, 3000, 15), (5, 700, 6), (32, 260, 20)])
def test_host_env_delta_transport_matches_oracle(cw, size, N, max_steps):
    """Delta transport: the device ships 16-byte records, the host library patches the caller's pinned frame mirror.
    After EVERY step the host frames must equal the oracle's full render (incl. re-seeded worlds and goal frames)."""
    seed, K = 91, 60
    env = cw.HostCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, seed=seed, transport="delta")
    ob = native.OracleBatch(native.make_config(H=size, W=size, max_steps=max_steps), N, seed=seed)
    obs = env.reset()
    o_goal = ob.reset(with_goal=True)
    o_obs = ob.render()
    assert np.array_equal(obs["observation"], o_obs) and np.array_equal(obs["desired_goal"], o_goal)
    rng = np.random.RandomState(2)
    new_goal = np.zeros_like(o_goal)
    for k in range(K):
        a = rng.randint(0, 6, N)
        obs, reward, done, _ = env.step(a)
        o_reward, o_done = ob.step_full(a.astype(np.uint8), auto_reset=True, obs=o_obs, goal_obs=new_goal)
        o_goal[o_done == 1] = new_goal[o_done == 1]
        assert np.array_equal(reward, o_reward) and np.array_equal(done, o_done.astype(bool)), k
        assert np.array_equal(obs["observation"], o_obs), f"host frame mirror diverged at step {k}"
        assert np.array_equal(obs["desired_goal"], o_goal), f"host goal mirror diverged at step {k}"
    assert np.array_equal(env.stats(), ob.stats) and ob.stats[0] > 0
    env.close()


def test_randomized_configurations_vs_oracle(cw):
    """Fuzz: random grid sizes (3..40), batch sizes (incl. non-multiples of the group size), episode lengths, task
    subsets, stacking, reward style and fixed pools; fused step + auto-reset + goal/init frames against the oracle."""
    rng = np.random.RandomState(20260101)
    names = cw.TASK_LIST
    for trial in range(36):
        size = int(rng.choice([3, 4, 5, 6, 7, 9, 11, 13, 16, 21, 24, 33, 40]))
        N = int(rng.randint(1, 400))
        max_steps = int(rng.randint(1, 14))
        nsel = int(rng.randint(1, 10))
        sel_idx = sorted(rng.choice(9, nsel, replace=False).tolist())
        ntasks = int(rng.randint(1, nsel + 1))
        stacking = bool(rng.randint(2))
        subset = bool(rng.randint(2))
        pool = int(rng.choice([0, 0, 3]))
        seed = int(rng.randint(1 << 30))
        base = int(rng.randint(1 << 20))
        env = cw.BatchedCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, selected_tasks=[names[i] for i in sel_idx],
                                         number_of_tasks=ntasks, stacking=stacking, reward_style=("s" if subset else None),
                                         fixed_init_state=pool, seed=seed, env_id_base=base)
        cfg = native.make_config(H=size, W=size, max_steps=max_steps, subset_reward=subset, stacking=stacking,
                                 selected=tuple(sel_idx), number_of_tasks=ntasks)
        ob = native.OracleBatch(cfg, N, seed=seed, env_id_base=base)
        if pool:                                       # the C oracle keeps raw pointers: the arrays must outlive the trial
            pool_g, pool_a = env._fixed_grid.cpu().numpy().copy(), env._fixed_agent.cpu().numpy().astype(np.uint32)
            native.set_fixed_pool(pool_g, pool_a)
        try:
            env.reset()
            o_goal = ob.reset(with_goal=True)
            o_obs = ob.render()
            new_goal = np.zeros_like(o_goal)
            K = int(rng.randint(5, 30))
            for k in range(K):
                a = rng.randint(0, 6, N).astype(np.uint8)
                _, reward, done, _ = env.step(a)
                o_reward, o_done = ob.step_full(a, auto_reset=True, obs=o_obs, goal_obs=new_goal)
                o_goal[o_done == 1] = new_goal[o_done == 1]
                where = f"trial {trial} size {size} N {N} max_steps {max_steps} step {k}"
                assert np.array_equal(reward.cpu().numpy(), o_reward) and np.array_equal(done.cpu().numpy(), o_done.astype(bool)), where
            assert_env_equals_oracle(env, ob, where)
            assert np.array_equal(env.obs.cpu().numpy(), o_obs), where
            assert np.array_equal(env.desired_goal.cpu().numpy(), o_goal), where
            assert np.array_equal(env.stats.cpu().numpy(), ob.stats), where
        finally:
            native.set_fixed_pool(None)


@pytest.mark.parametrize("size,N,max_steps", [(21, 3000, 15), (7, 130, 5), (32, 700, 20), (5, 1, 3), (6, 65, 4), (21, 900, 1), (9, 2100, 2),
                                               (21, 20000, 40)])
def test_host_env_device_consumer_matches_oracle(cw, size, N, max_steps):
    """Device-consumer transport (return_frames=False): reward / done land in mapped host memory as one status byte per world
    and the call returns without a stream synchronisation, frames stay in HBM (rotating buffers).  Up to 16 384 worlds a step is
    the two-launch pipeline (step launch + render launch of its snapshot; 1- and 2-step episodes re-seed a world in consecutive
    steps), above that one fused chained launch.  reward / done after every call and the device frames (fetched) every few
    steps must equal the oracle's."""
    seed, K = 17, 70
    env = cw.HostCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, seed=seed, return_frames=False)
    ob = native.OracleBatch(native.make_config(H=size, W=size, max_steps=max_steps), N, seed=seed)
    env.reset()
    o_goal = ob.reset(with_goal=True)
    o_obs = ob.render()
    obs, goal = env.fetch_frames()
    assert np.array_equal(obs, o_obs) and np.array_equal(goal, o_goal)
    rng = np.random.RandomState(3)
    new_goal = np.zeros_like(o_goal)
    for k in range(K):
        a = rng.randint(0, 6, N)
        _, reward, done, _ = env.step(a)
        o_reward, o_done = ob.step_full(a.astype(np.uint8), auto_reset=True, obs=o_obs, goal_obs=new_goal)
        o_goal[o_done == 1] = new_goal[o_done == 1]
        assert np.array_equal(reward, o_reward) and np.array_equal(done, o_done.astype(bool)), k
        if k % 7 == 6 or k == K - 1:
            obs, goal = env.fetch_frames()
            assert np.array_equal(obs, o_obs), f"device frames diverged at step {k}"
            assert np.array_equal(goal, o_goal), f"device goal frames diverged at step {k}"
    assert np.array_equal(env.stats(), ob.stats) and ob.stats[0] > 0
    env.close()


@pytest.mark.parametrize("frames", [False, True])
def test_host_env_step_many_matches_oracle(cw, frames):
    """cw_host_step_many: K steps of an open-loop tape per call (device consumer: K chained launches, one wait; more steps than
    notification slots in one call; host frames by delta: K single steps)."""
    N, size, max_steps, seed = 1500, 21, 12, 29
    env = cw.HostCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, seed=seed, return_frames=frames, transport="delta")
    ob = native.OracleBatch(native.make_config(H=size, W=size, max_steps=max_steps), N, seed=seed)
    env.reset(); ob.reset(with_goal=True)
    o_obs = ob.render()
    rng = np.random.RandomState(4)
    for K in (1, 5, 130, 64):
        tape = rng.randint(0, 6, (K, N)).astype(np.uint8)
        obs, reward, done, _ = env.step_many(tape)
        for k in range(K):
            o_reward, o_done = ob.step_full(tape[k], auto_reset=True, obs=o_obs)
            assert np.array_equal(reward[k], o_reward) and np.array_equal(done[k], o_done.astype(bool)), (K, k)
        frame = obs["observation"] if frames else env.fetch_frames()[0]
        assert np.array_equal(frame, o_obs), K
    _, reward, done, _ = env.step(tape[0])                       # single steps and runs interleave freely
    o_reward, o_done = ob.step_full(tape[0], auto_reset=True, obs=o_obs)
    assert np.array_equal(reward, o_reward)
    assert np.array_equal(env.stats(), ob.stats)
    env.close()


def test_host_api_unbound_pageable_buffers_and_delta_rules(cw):
    """C ABI directly: arrays that were never declared with cw_host_bind (pageable NumPy memory, new ones every call) are staged
    by the library; a delta handle refuses a step without a frame mirror; a different mirror pointer triggers a full refresh;
    cw_host_load_state injects reference-style states."""
    import ctypes as C
    from gym_craftingworld_b200 import _lib
    lib = _lib.load()
    N, size, seed = 600, 9, 5
    cfg = cw.make_config(size=(size, size), max_steps=11)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.RandomState(8)
    for flags in (_lib.F_AUTO_RESET, _lib.F_AUTO_RESET | _lib.F_DELTA_TRANSPORT):
        ob = native.OracleBatch(native.make_config(H=size, W=size, max_steps=11), N, seed=seed)
        h = C.c_void_p()
        _lib.check(lib.cw_host_create(C.byref(cfg), N, 0, C.c_uint64(seed), C.c_uint64(0), flags, C.byref(h)))
        delta = bool(flags & _lib.F_DELTA_TRANSPORT)
        frames = np.zeros((N, 4 * size, 4 * size, 3), np.uint8)
        _lib.check(lib.cw_host_reset(h, p(frames), None))
        ob.reset(); o_obs = ob.render()
        assert np.array_equal(frames, o_obs)
        for k in range(25):
            a = rng.randint(0, 6, N).astype(np.uint8).copy()      # fresh pageable arrays every call
            rew, dn = np.zeros(N, np.int32), np.zeros(N, np.uint8)
            if delta and k == 10:
                frames = np.zeros_like(frames)                     # a NEW mirror: the library must refresh it in full
            if delta and k == 5:
                assert lib.cw_host_step(h, p(a), p(rew), p(dn), None) == -1   # CW_E_BADCONFIG
            _lib.check(lib.cw_host_step(h, p(a), p(rew), p(dn), p(frames) if (delta or k % 2) else None))
            o_rew, o_dn = ob.step_full(a, auto_reset=True, obs=o_obs)
            assert np.array_equal(rew, o_rew) and np.array_equal(dn, o_dn), (flags, k)
            if delta or k % 2:
                assert np.array_equal(frames, o_obs), (flags, k)
        # inject: every world gets world 0's grid / agent, step counters staggered
        g = np.repeat(ob.grid[:1], N, 0); ag = np.repeat(ob.agent[:1], N, 0); gl = np.repeat(ob.goal[:1], N, 0)
        t = (np.arange(N) % 11).astype(np.int32)
        _lib.check(lib.cw_host_load_state(h, p(g), p(ag), p(gl), p(t), p(frames)))
        ob.grid[:] = g; ob.init_grid[:] = g; ob.agent[:] = ag; ob.goal[:] = gl; ob.t[:] = t
        o_obs = ob.render()
        assert np.array_equal(frames, o_obs)
        a = rng.randint(0, 6, N).astype(np.uint8)
        rew, dn = np.zeros(N, np.int32), np.zeros(N, np.uint8)
        _lib.check(lib.cw_host_step(h, p(a), p(rew), p(dn), p(frames)))
        o_rew, o_dn = ob.step_full(a, auto_reset=True, obs=o_obs)
        assert np.array_equal(rew, o_rew) and np.array_equal(dn, o_dn) and np.array_equal(frames, o_obs)
        _lib.check(lib.cw_host_destroy(h))


@pytest.mark.parametrize("max_steps", [1, 2, 7])
def test_compact_step_with_predrawn_reset_records_matches_oracle(cw, max_steps):
    """cw_step / cw_rollout with pre-drawn reset records (CwState.reset_rec / reset_list): finished worlds are re-seeded by a copy
    of a record drawn ahead of time by refill CTAs.  Episodes of 1 or 2 steps make every world finish (almost) every step, so
    records are consumed while their successors are still being drawn -- the stale-tag fallback and the fast path must both
    give the oracle's bits, step by step and inside multi-step rollouts, and after a change of seed."""
    N, size, seed = 5000, 6, 31
    env = cw.BatchedCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, seed=seed, obs_mode="compact")
    assert env.reset_rec is not None
    ob = native.OracleBatch(native.make_config(H=size, W=size, max_steps=max_steps), N, seed=seed)
    env.reset(); ob.reset()
    rng = np.random.RandomState(6)
    for k in range(40):
        a = rng.randint(0, 6, N).astype(np.uint8)
        _, reward, done, _ = env.step(torch.from_numpy(a).cuda())
        o_reward, o_done = ob.step_full(a, auto_reset=True)
        assert np.array_equal(reward.cpu().numpy(), o_reward) and np.array_equal(done.cpu().numpy(), o_done.astype(bool)), k
    assert_env_equals_oracle(env, ob, f"single steps, max_steps {max_steps}")
    if max_steps == 7:                                             # the records run one episode ahead of (nearly) every world
        tags = env.reset_rec[:, 0].cpu().numpy().astype(np.uint32)
        assert (tags == env.episode.cpu().numpy().astype(np.uint32)).mean() > 0.7
    tape = rng.randint(0, 6, (33, N)).astype(np.uint8)            # several finishes per world inside ONE launch
    rew, dn = env.rollout(torch.from_numpy(tape).cuda())
    for k in range(33):
        o_reward, o_done = ob.step_full(tape[k], auto_reset=True)
        assert np.array_equal(rew[k].cpu().numpy(), o_reward) and np.array_equal(dn[k].cpu().numpy(), o_done.astype(bool)), k
    assert_env_equals_oracle(env, ob, f"rollout, max_steps {max_steps}")
    assert np.array_equal(env.stats.cpu().numpy(), ob.stats)
    # a new key: the records drawn under the old one must not be used
    env.seed(seed + 1)
    ob2 = native.OracleBatch(native.make_config(H=size, W=size, max_steps=max_steps), N, seed=seed + 1)
    ob2.grid[:] = ob.grid; ob2.init_grid[:] = ob.init_grid; ob2.agent[:] = ob.agent; ob2.goal[:] = ob.goal; ob2.t[:] = ob.t
    ob2.episode[:] = ob.episode
    for k in range(10):
        a = rng.randint(0, 6, N).astype(np.uint8)
        env.step(torch.from_numpy(a).cuda())
        ob2.step_full(a, auto_reset=True)
    assert_env_equals_oracle(env, ob2, "after seed()")


def test_host_env_device_consumer_survives_chain_wrap(cw):
    """More single steps than a chain has positions (CW_CHAIN_MAX_POS = 1024): the handle re-opens the chain and carries on."""
    N, size, max_steps, seed = 96, 5, 9, 41
    env = cw.HostCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, seed=seed, return_frames=False)
    ob = native.OracleBatch(native.make_config(H=size, W=size, max_steps=max_steps), N, seed=seed)
    env.reset(); ob.reset()
    o_obs = ob.render()
    rng = np.random.RandomState(9)
    acts = rng.randint(0, 6, (1100, N)).astype(np.uint8)
    for k in range(1100):
        _, reward, done, _ = env.step(acts[k])
        o_reward, o_done = ob.step_full(acts[k], auto_reset=True, obs=o_obs)
        assert np.array_equal(reward, o_reward) and np.array_equal(done, o_done.astype(bool)), k
    obs, _, _, _ = env.step_many(acts[:300])                       # and an open-loop run across the next wrap
    for k in range(300):
        ob.step_full(acts[k], auto_reset=True, obs=o_obs)
    assert np.array_equal(env.fetch_frames()[0], o_obs)
    assert np.array_equal(env.stats(), ob.stats)
    env.close()
