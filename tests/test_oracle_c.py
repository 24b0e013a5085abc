"""CPU: the C restatement (oracle/cw_oracle.c) against the frozen reference traces, the Philox known-answer
vectors, and the NumPy spec (oracle/compact.py) for the Philox reset / imagine_obs paths."""
import numpy as np
import pytest

from oracle import compact, native
from tests import golden_util as gu


def batch_from_golden(d, **kw):
    cfg = native.make_config(H=d["H"], W=d["W"], max_steps=d["max_steps"], subset_reward=bool(d["subset"]))
    B = d["actions"].shape[0]
    ob = native.OracleBatch(cfg, B, **kw)
    ob.load_state(d["grid0"], d["r0"], d["c0"], d["hold0"], d["desired"])
    return ob


@pytest.mark.parametrize("name", gu.golden_files())
def test_c_oracle_matches_reference_trace(name, oracle_lib):
    d = gu.load(name)
    ob = batch_from_golden(d)
    B, T = d["actions"].shape
    fidx = {int(t): i for i, t in enumerate(d["frame_t"])}
    assert np.array_equal(ob.render(), d["frame0"])
    for t in range(T):
        reward, done = ob.step(d["actions"][:, t])
        where = f"{name} step {t}"
        assert np.array_equal(ob.grid2d, d["grid"][:, t]), where
        assert np.array_equal(ob.r, d["r"][:, t]) and np.array_equal(ob.c, d["c"][:, t]), where
        assert np.array_equal(ob.hold, d["hold"][:, t]), where
        assert np.array_equal(ob.achieved, d["achieved"][:, t]), where
        assert np.array_equal(reward, d["reward"][:, t]) and np.array_equal(done, d["done"][:, t]), where
        frames = ob.render()
        assert [gu.crc(f) for f in frames] == list(d["frame_crc"][:, t]), where
        if t in fidx:
            assert np.array_equal(frames, d["frames"][:, fidx[t]]), where


# Random123 known-answer vectors for philox4x32-10 (kat_vectors of the Random123 distribution)
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,expect", PHILOX_KAT)
def test_philox_known_answers(ctr, key, expect, oracle_lib):
    assert compact.philox4x32_10(ctr, key) == expect
    assert native.philox(ctr, key) == expect


def test_philox_stream_uniform_c_equals_numpy(oracle_lib):
    for n in (1, 2, 9, 441, 1024, 3_000_000_000):
        s = compact.PhiloxStream(1234567890123, 42, 7)
        want = [s.uniform(n) for _ in range(50)]
        got = native.stream_uniform(1234567890123, 42, 7, n, 50)
        assert list(got) == want
        assert max(want) < n


@pytest.mark.parametrize("H,W,stacking,selected,ntasks", [
    (21, 21, True, tuple(range(9)), None),
    (32, 32, True, (3, 2), 2),
    (8, 8, False, (0, 1, 5, 8), None),
    (4, 4, True, tuple(range(9)), 3),
])
def test_reset_and_imagine_c_equals_numpy(H, W, stacking, selected, ntasks, oracle_lib):
    cfg = native.make_config(H=H, W=W, stacking=stacking, selected=selected, number_of_tasks=ntasks)
    pc = compact.Config(H=H, W=W, stacking=stacking, selected=selected,
                        number_of_tasks=ntasks if ntasks is not None else len(selected))
    N, seed, base = 24, 99, 1000
    ob = native.OracleBatch(cfg, N, seed=seed, env_id_base=base)
    for episode in range(3):
        goal_obs = ob.reset(with_goal=True)
        for n in range(N):
            s, (ig, ir, ic, ih) = compact.reset_env(seed, base + n, episode, pc, with_goal=True)
            assert np.array_equal(ob.grid2d[n], s.grid) and np.array_equal(ob.init_grid[n], ob.grid[n])
            assert (ob.r[n], ob.c[n], ob.hold[n]) == (s.r, s.c, 0)
            assert ob.desired[n] == s.desired and ob.achieved[n] == 0 and ob.t[n] == 0
            assert ob.episode[n] == episode + 1
            assert np.array_equal(goal_obs[n], compact.render(ig, ir, ic, ih))
        assert (ob.grid2d > 0).sum() == 8 * N                       # one of each object
        assert all(ob.grid2d[n, ob.r[n], ob.c[n]] == 0 for n in range(N))   # agent starts on an empty cell


def test_imagine_on_dense_worlds_c_equals_numpy(oracle_lib):
    d = gu.load("dense_8x8.npz")
    ob = batch_from_golden(d, seed=5, env_id_base=77)
    out_grid, out_agent = ob.imagine()
    for n in range(ob.N):
        rng = compact.PhiloxStream(5, 77 + n, 0)
        g, r, c, h = compact.imagine(d["grid0"][n], int(d["r0"][n]), int(d["c0"][n]), int(d["hold0"][n]),
                                     int(d["desired"][n]), rng)
        assert np.array_equal(out_grid[n, :64].reshape(8, 8), g), n
        assert out_agent[n] == (r | (c << 8) | (h << 16)), n


def test_masked_reset_only_touches_masked(oracle_lib):
    cfg = native.make_config()
    ob = native.OracleBatch(cfg, 16, seed=3)
    ob.reset()
    before = (ob.grid.copy(), ob.agent.copy(), ob.goal.copy(), ob.episode.copy())
    mask = np.zeros(16, np.uint8)
    mask[[2, 5, 11]] = 1
    ob.t[:] = 17
    ob.reset(mask=mask)
    keep = mask == 0
    assert np.array_equal(ob.grid[keep], before[0][keep]) and np.array_equal(ob.agent[keep], before[1][keep])
    assert (ob.t[keep] == 17).all() and (ob.t[~keep] == 0).all()
    assert np.array_equal(ob.episode, before[3] + mask)
    assert not np.array_equal(ob.grid[~keep], before[0][~keep])


def test_step_full_autoreset_and_stats(oracle_lib):
    """auto-reset: reward/done of the finished episode are returned, state is the new episode's, stats add up."""
    cfg = native.make_config(H=5, W=5, max_steps=10)
    N = 64
    ob = native.OracleBatch(cfg, N, seed=11)
    ob.reset()
    rng = np.random.RandomState(0)
    obs = np.zeros(ob.frame_shape(), np.uint8)
    ep_done = 0
    ret = 0
    for k in range(200):
        a = rng.randint(0, 6, N)
        t_before = ob.t.copy()
        reward, done = ob.step_full(a, auto_reset=True, obs=obs)
        ep_done += int(done.sum())
        ret += int(reward.sum())
        assert ((ob.t == 0) == (done == 1)).all()
        assert ((done == 1) == ((t_before + 1 >= 10) | (reward == 10))).all()
        assert np.array_equal(obs, ob.render())
    assert ob.stats[0] == ep_done and ob.stats[1] == (ob.stats[2] + ob.stats[3]) // 11
    # return_sum over finished episodes + partial returns of running ones == all rewards
    assert ob.stats[2] - int(ob.t.sum()) == ret
    assert ob.stats[3] + int(ob.t.sum()) == 200 * N


def test_run_threads_equals_step_full(oracle_lib):
    cfg = native.make_config(H=8, W=8, max_steps=20)
    N, K = 96, 50
    acts = np.random.RandomState(1).randint(0, 6, (K, N)).astype(np.uint8)
    a = native.OracleBatch(cfg, N, seed=2)
    a.reset()
    obs_a = np.zeros(a.frame_shape(), np.uint8)
    for k in range(K):
        a.step_full(acts[k], obs=obs_a)
    for mode in (1, 2):
        b = native.OracleBatch(cfg, N, seed=2)
        b.reset()
        obs_b = b.render()
        b.run_threads(acts, render_mode=mode, obs=obs_b, nthreads=3)
        assert np.array_equal(a.grid, b.grid) and np.array_equal(a.agent, b.agent) and np.array_equal(a.goal, b.goal)
        assert np.array_equal(a.stats, b.stats)
        assert np.array_equal(obs_a, obs_b), f"render mode {mode}"      # incremental render_edit == full render
