"""GPU (-m gpu): parity at BASELINE.json's full sizes.  Where the C oracle finishes in seconds it checks everything
bit-for-bit; at the largest sizes the state is still checked bit-for-bit against the oracle and the pixels through
size-independent properties (render idempotence across the two GPU paths + an oracle-rendered sample of worlds)."""
import numpy as np
import pytest
import torch

from oracle import native

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cw():
    import gym_craftingworld_b200 as pkg
    assert torch.cuda.is_available()
    return pkg


def oracle_like(env, seed):
    cfg = native.make_config(H=env.cfg.H, W=env.cfg.W, max_steps=env.cfg.max_steps)
    return native.OracleBatch(cfg, env.num_envs, seed=seed, env_id_base=env.env_id_base)


def state_equal(env, ob):
    return (np.array_equal(env.grid.cpu().numpy(), ob.grid) and np.array_equal(env.init_grid.cpu().numpy(), ob.init_grid)
            and np.array_equal(env.agent.cpu().numpy().astype(np.uint32), ob.agent)
            and np.array_equal(env.goal.cpu().numpy().astype(np.uint32), ob.goal)
            and np.array_equal(env.t.cpu().numpy(), ob.t)
            and np.array_equal(env.episode.cpu().numpy().astype(np.uint32), ob.episode))


def dense_worlds(rng, N, H, W):
    """BASELINE config 5 placement: p=0.5 occupancy, uniform types, >=1 of each, agent on an empty cell."""
    grid = np.where(rng.random_sample((N, H * W)) < 0.5, rng.randint(1, 9, (N, H * W)), 0).astype(np.uint8)
    cells = np.argsort(rng.random_sample((N, H * W)), axis=1)[:, :9]
    rows = np.arange(N)
    for k in range(8):
        grid[rows, cells[:, k]] = k + 1
    grid[rows, cells[:, 8]] = 0
    return grid.reshape(N, H, W), (cells[:, 8] // W).astype(np.uint8), (cells[:, 8] % W).astype(np.uint8)


def test_config2_4096_worlds_full_episode(cw):
    """Config 2: 4096 worlds, default grid, nine-skill tasks, auto-reset -- 330 fused steps (every world times out and
    re-seeds at least once) against the oracle: reward/done every step; state + all three frame buffers at checkpoints."""
    N, seed, K = 4096, 2, 330
    env = cw.BatchedCraftingWorldEnv(N, seed=seed, auto_reset=True)
    ob = oracle_like(env, seed)
    env.reset()
    o_goal = ob.reset(with_goal=True)
    o_obs = ob.render()
    o_init = o_obs.copy()
    acts = torch.randint(0, 6, (K, N), generator=torch.Generator().manual_seed(5), dtype=torch.uint8)
    acts_np, acts_gpu = acts.numpy(), acts.cuda()
    new_goal = np.zeros_like(o_goal)
    for k in range(K):
        obs, reward, done, _ = env.step(acts_gpu[k])
        o_reward, o_done = ob.step_full(acts_np[k], auto_reset=True, obs=o_obs, goal_obs=new_goal)
        fresh = o_done == 1
        if fresh.any():
            o_goal[fresh] = new_goal[fresh]
            o_init[fresh] = o_obs[fresh]
        assert np.array_equal(reward.cpu().numpy(), o_reward) and np.array_equal(done.cpu().numpy(), fresh), f"step {k}"
        if k in (0, 150, 298, 299, 300, K - 1):
            assert state_equal(env, ob), f"state at step {k}"
            assert np.array_equal(obs["observation"].cpu().numpy(), o_obs), f"frames at step {k}"
            assert np.array_equal(obs["desired_goal"].cpu().numpy(), o_goal), f"goal frames at step {k}"
            assert np.array_equal(obs["init_observation"].cpu().numpy(), o_init), f"init frames at step {k}"
    assert np.array_equal(env.stats.cpu().numpy(), ob.stats) and ob.stats[0] >= N


def test_config2_chained_graph_full_episode(cw):
    """Config 2 exactly as bench.py runs it: 330 steps captured as ONE chain of launches (cw_step_render_chained) in a CUDA
    graph, frames rotating over 4 buffers -- final state, statistics, goal / init frames and the last 4 frames of every world
    against the oracle."""
    N, seed, K, R = 4096, 2, 330, 4
    env = cw.BatchedCraftingWorldEnv(N, seed=seed, auto_reset=True, obs_buffers=R)
    ob = oracle_like(env, seed)
    env.reset()
    o_goal = ob.reset(with_goal=True)
    o_obs = ob.render()
    o_init = o_obs.copy()
    acts = torch.randint(0, 6, (K, N), generator=torch.Generator().manual_seed(5), dtype=torch.uint8)
    acts_np, acts_gpu = acts.numpy(), acts.cuda()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for k in range(K):
                env.step(acts_gpu[k], chain_pos=k)
        g.replay()
    new_goal = np.zeros_like(o_goal)
    last = {}
    for k in range(K):
        _, o_done = ob.step_full(acts_np[k], auto_reset=True, obs=o_obs, goal_obs=new_goal)
        fresh = o_done == 1
        if fresh.any():
            o_goal[fresh] = new_goal[fresh]
            o_init[fresh] = o_obs[fresh]
        if k >= K - R:
            last[k] = o_obs.copy()
    torch.cuda.synchronize()
    assert state_equal(env, ob) and np.array_equal(env.stats.cpu().numpy(), ob.stats) and ob.stats[0] >= N
    assert np.array_equal(env.desired_goal.cpu().numpy(), o_goal) and np.array_equal(env.init_obs.cpu().numpy(), o_init)
    for k in range(K - R, K):                                      # step k wrote ring buffer (k + 1) % R
        assert np.array_equal(env._obs_ring[(k + 1) % R].cpu().numpy(), last[k]), f"frames of step {k}"


def test_config4_chained_131072_worlds(cw):
    """Config 4 per-GPU slice through chained launches (persistent CTAs, 2 frame buffers): state + statistics bit-exact against
    the oracle, every pixel of the last two steps against a fresh render of the corresponding state."""
    N, seed, K = 131072, 6, 12
    env = cw.BatchedCraftingWorldEnv(N, max_steps=10, seed=seed, auto_reset=True, goal_images=False, obs_buffers=2)
    ref = cw.BatchedCraftingWorldEnv(N, max_steps=10, seed=seed, auto_reset=True, goal_images=False)
    ob = oracle_like(env, seed)
    env.reset(); ref.reset(); ob.reset()
    acts = np.random.RandomState(9).randint(0, 6, (K, N)).astype(np.uint8)
    a_gpu = torch.from_numpy(acts).cuda()
    for k in range(K):
        env.step(a_gpu[k], chain_pos=k)
        ob.step_full(acts[k], auto_reset=True)
    for k in range(K - 1):
        ref.step(a_gpu[k])
    assert torch.equal(env._obs_ring[(K - 1) % 2], ref.obs)         # step K-2
    ref.step(a_gpu[K - 1])
    assert torch.equal(env._obs_ring[K % 2], ref.obs)               # step K-1
    assert state_equal(env, ob) and np.array_equal(env.stats.cpu().numpy(), ob.stats) and ob.stats[0] > N


def test_config5_16384_dense_32x32(cw):
    """Config 5: 16384 dense 32x32 worlds -- 24 fused steps; final state and EVERY pixel (805 MB) against the oracle."""
    N, H, seed, K = 16384, 32, 3, 24
    env = cw.BatchedCraftingWorldEnv(N, size=(H, H), seed=seed, auto_reset=True, goal_images=False)
    ob = oracle_like(env, seed)
    rng = np.random.RandomState(11)
    grid, r, c = dense_worlds(rng, N, H, H)
    desired = rng.randint(1, 512, N)
    env.reset(); ob.reset()                                        # episode counters advance identically
    env.load_state(grid, r, c, np.zeros(N), desired)
    ob.load_state(grid, r, c, np.zeros(N), desired)
    acts = rng.randint(0, 6, (K, N)).astype(np.uint8)
    acts_gpu = torch.from_numpy(acts).cuda()
    for k in range(K):
        _, reward, done, _ = env.step(acts_gpu[k])
        o_reward, o_done = ob.step_full(acts[k], auto_reset=True)
        assert np.array_equal(reward.cpu().numpy(), o_reward) and np.array_equal(done.cpu().numpy(), o_done.astype(bool)), k
    assert state_equal(env, ob)
    assert (env.grid.cpu().numpy() > 0).mean() > 0.4               # still dense
    frames = env.obs.cpu().numpy()
    assert np.array_equal(frames, ob.render())
    del frames


def test_config3_65536_worlds_compact_rollout(cw):
    """Config 3: 65536 worlds, compact observations: single steps and a 96-step rollout launch against the oracle."""
    N, seed, K = 65536, 4, 96
    acts = np.random.RandomState(8).randint(0, 6, (K, N)).astype(np.uint8)
    env = cw.BatchedCraftingWorldEnv(N, max_steps=40, seed=seed, obs_mode="compact")
    ob = oracle_like(env, seed)
    env.reset(); ob.reset()
    o_rew = np.zeros((K, N), np.int32); o_dn = np.zeros((K, N), np.uint8)
    for k in range(K):
        o_rew[k], o_dn[k] = ob.step_full(acts[k], auto_reset=True)
    rew, dn = env.rollout(acts)
    assert np.array_equal(rew.cpu().numpy(), o_rew) and np.array_equal(dn.cpu().numpy(), o_dn.astype(bool))
    assert state_equal(env, ob) and np.array_equal(env.stats.cpu().numpy(), ob.stats)
    env2 = cw.BatchedCraftingWorldEnv(N, max_steps=40, seed=seed, obs_mode="compact")
    env2.reset()
    a_gpu = torch.from_numpy(acts).cuda()
    for k in range(K):
        env2.step(a_gpu[k])
    assert state_equal(env2, ob) and np.array_equal(env2.stats.cpu().numpy(), ob.stats)


def test_config4_131072_worlds_properties(cw):
    """Config 4 per-GPU slice (131072 worlds, 2.8 GB of frames per step): state bit-exact against the oracle; pixels via
    (i) idempotence: the fused kernel's frames == a fresh cw_render of the same state, all 2.8 GB, compared on the
    device, and (ii) an oracle-rendered sample of 4096 worlds spread over the batch."""
    N, seed, K = 131072, 6, 12
    env = cw.BatchedCraftingWorldEnv(N, max_steps=10, seed=seed, auto_reset=True, goal_images=False)
    ob = oracle_like(env, seed)
    env.reset(); ob.reset()
    acts = np.random.RandomState(9).randint(0, 6, (K, N)).astype(np.uint8)
    a_gpu = torch.from_numpy(acts).cuda()
    for k in range(K):
        _, reward, done, _ = env.step(a_gpu[k])
        o_reward, o_done = ob.step_full(acts[k], auto_reset=True)
        assert np.array_equal(reward.cpu().numpy(), o_reward) and np.array_equal(done.cpu().numpy(), o_done.astype(bool)), k
    assert state_equal(env, ob) and np.array_equal(env.stats.cpu().numpy(), ob.stats) and ob.stats[0] > N
    fused = env.obs.clone()
    again = env.render()                                            # render-only launch into the same buffer
    assert torch.equal(fused, again)
    idx = np.linspace(0, N - 1, 4096).astype(np.int64)
    sub = native.OracleBatch(ob.cfg, len(idx))
    sub.grid[:], sub.agent[:] = ob.grid[idx], ob.agent[idx]
    assert np.array_equal(fused[torch.from_numpy(idx).cuda()].cpu().numpy(), sub.render())
    # structural invariants: agent inside the grid, held item in 0..3, reward in {-1, MAX_STEPS}
    ag = env.agent.cpu().numpy()
    assert ((ag & 0xFF) < 21).all() and (((ag >> 8) & 0xFF) < 21).all() and (((ag >> 16) & 0xFF) <= 3).all()
    assert set(np.unique(env.reward.cpu().numpy())) <= {-1, 10}


def test_bench_line_contract():
    """`python bench.py` (our arm, shortened): ONE JSON line with the contract's keys, a roofline record that can be recomputed
    from the line, declared e2e copy sizes, and the same `config` dict the reference arm prints."""
    import argparse
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "20", "--warmup", "5", "--only", "--quick", "--no-cpu-baseline", "--no-steady",
                          "--no-incremental", "--no-closed-loop"], cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "method", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["steps"] == 20 and d["gpu_launches"] == 20 and d["n_gpus"] == 1 and d["dtype"] == "u8" and d["vs_baseline"] is None
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    algorithmic = r["algorithmic_bytes_per_env_step"] * r["units_per_launch"]
    assert abs(r["achieved"] - algorithmic / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert abs(d["value"] - 4096 * 20 / (d["ms_per_step"] * 20e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 4096 and e["d2h_bytes_per_step"] > 0
    sys.path.insert(0, root)
    import bench
    assert d["config"] == bench.workload_config(argparse.Namespace(workload="cfg2", envs=0, ring=0), 1)
