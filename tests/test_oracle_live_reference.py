"""CPU, builder container only: live differential tests of the oracle against the UNMODIFIED reference
(skipped where /root/reference is absent, e.g. on the GPU box, where the frozen tests/golden traces stand in)."""
import collections

import numpy as np
import pytest

from oracle import compact, native, ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference source tree not present")


@pytest.fixture(scope="module")
def ray():
    return ref_shim.load_reference()


def random_world(rng, H, W):
    density = rng.choice([0.15, 0.4, 0.8])
    grid = np.where(rng.random_sample((H, W)) < density, rng.randint(1, 9, (H, W)), 0).astype(np.uint8)
    r, c = int(rng.randint(H)), int(rng.randint(W))
    if rng.random_sample() < 0.6:
        grid[r, c] = 0
    return grid, r, c, int(rng.choice([0, 1, 2, 3])), int(rng.randint(1, 512))


@pytest.mark.parametrize("size,subset", [(4, False), (6, True), (7, False), (21, False)])
def test_live_differential_step_render(ray, size, subset, oracle_lib):
    """Fresh random worlds (not the frozen ones): reference vs C oracle vs NumPy spec, every field, every pixel."""
    rng = np.random.RandomState(1000 + size)
    B, T = (24, 60) if size < 21 else (6, 80)
    worlds = [random_world(rng, size, size) for _ in range(B)]
    cfgc = native.make_config(H=size, W=size, max_steps=40, subset_reward=subset)
    pc = compact.Config(H=size, W=size, max_steps=40, subset_reward=subset)
    ob = native.OracleBatch(cfgc, B)
    ob.load_state(np.stack([w[0] for w in worlds]), [w[1] for w in worlds], [w[2] for w in worlds],
                  [w[3] for w in worlds], [w[4] for w in worlds])
    envs = [ref_shim.make_injected_env(ray, *w, max_steps=40, reward_style=("s" if subset else None)) for w in worlds]
    pys = [compact.EnvState(w[0].copy(), w[0].copy(), w[1], w[2], w[3], 0, w[4]) for w in worlds]
    for t in range(T):
        a = rng.randint(0, 6, B)
        reward, done = ob.step(a)
        frames = ob.render()
        for b, env in enumerate(envs):
            _, rw, dn, info = env.step(int(a[b]))
            g, r, c, h, ach, px = ref_shim.read_back(env)
            assert np.array_equal(ob.grid2d[b], g) and (ob.r[b], ob.c[b], ob.hold[b]) == (r, c, h)
            assert ob.achieved[b] == ach and reward[b] == rw and bool(done[b]) == dn
            assert np.array_equal(frames[b], px)
            assert np.array_equal(px, env.render(env.obs_one_hot).astype(np.uint8))    # incremental == full upstream
            prw, pdn, _ = compact.step_env(pys[b], int(a[b]), pc)
            assert (prw, pdn, pys[b].achieved) == (rw, dn, ach) and np.array_equal(pys[b].grid, g)


def _chi2_two_sample(a, b):
    """Pearson chi-square statistic and dof for two count vectors over the same categories."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    keep = (a + b) > 0
    a, b = a[keep], b[keep]
    k1, k2 = np.sqrt(b.sum() / a.sum()), np.sqrt(a.sum() / b.sum())
    return float((((k1 * a - k2 * b) ** 2) / (a + b)).sum()), int(keep.sum() - 1)


def test_reset_distribution_matches_reference(ray, oracle_lib):
    """Philox reset vs the reference's RandomState reset: same distribution of (object -> cell) placements and of
    the number / identity of desired tasks (two-sample chi-square, far below the 1e-4 rejection threshold)."""
    from scipy.stats import chi2
    size, n = 5, 4000
    env = ray.CraftingWorldEnvRay(size=(size, size))
    env.seed(123)
    ref_cells = np.zeros((9, size * size))
    ref_ntasks, ref_task = np.zeros(10), np.zeros(9)
    for _ in range(n):
        env.reset()
        g, r, c, h, _, _ = ref_shim.read_back(env)
        for k in range(8):
            ref_cells[k, int(np.flatnonzero(g.reshape(-1) == k + 1)[0])] += 1
        ref_cells[8, r * size + c] += 1
        d = env.desired_goal_vector[0]
        ref_ntasks[int(d.sum())] += 1
        ref_task += d
    ob = native.OracleBatch(native.make_config(H=size, W=size), n, seed=321)
    ob.reset()
    our_cells = np.zeros((9, size * size))
    for k in range(8):
        our_cells[k] = np.bincount(np.argmax(ob.grid[:, :size * size] == k + 1, axis=1), minlength=size * size)
    our_cells[8] = np.bincount(ob.r.astype(int) * size + ob.c, minlength=size * size)
    des = ob.desired.astype(int)
    our_ntasks = np.bincount([bin(x).count("1") for x in des], minlength=10)
    our_task = np.array([((des >> i) & 1).sum() for i in range(9)])
    for k in range(9):
        stat, dof = _chi2_two_sample(ref_cells[k], our_cells[k])
        assert chi2.sf(stat, dof) > 1e-4, f"object {k} placement distribution differs"
    stat, dof = _chi2_two_sample(ref_ntasks, our_ntasks)
    assert chi2.sf(stat, dof) > 1e-4
    stat, dof = _chi2_two_sample(ref_task, our_task)
    assert chi2.sf(stat, dof) > 1e-4


def test_imagine_distribution_matches_reference(ray, oracle_lib):
    """imagine_obs: on one fixed world, for goals with a small outcome space the set of reachable imagined frames
    is identical and their frequencies agree (chi-square) between the reference RNG and the Philox stream; for the
    all-skills goal (thousands of outcomes) the per-pixel-cell colour marginals agree."""
    from scipy.stats import chi2
    size, n = 4, 4000
    env = ray.CraftingWorldEnvRay(size=(size, size))
    env.seed(5)
    env.reset()
    g0, r0, c0, h0, _, _ = ref_shim.read_back(env)

    def sample(desired):
        env.desired_goal_vector = ref_shim.mask_to_bits(desired).reshape(1, 9)
        ref = [env.imagine_obs().astype(np.uint8) for _ in range(n)]
        ours = []
        for ep in range(n):
            gi, ri, ci, hi = compact.imagine(g0, r0, c0, h0, desired, compact.PhiloxStream(9, 0, ep))
            ours.append(compact.render(gi, ri, ci, hi))
        return ref, ours

    small_goals = [1 << 8, 1 << 6, 1 << 7, 0b000000011, 0b000100100, 0b100001000, 0b000101100, 0b010010000]
    for desired in small_goals:
        ref, ours = sample(desired)
        cr = collections.Counter(x.tobytes() for x in ref)
        co = collections.Counter(x.tobytes() for x in ours)
        assert set(cr) == set(co), f"goal {desired:09b}: reachable imagined frames differ"
        keys = sorted(cr)
        if len(keys) > 1:
            stat, dof = _chi2_two_sample([cr[k] for k in keys], [co[k] for k in keys])
            assert chi2.sf(stat, dof) > 1e-4, f"imagine distribution differs for goal {desired:09b}"
    ref, ours = sample(0b111111111)
    mr = np.mean(np.stack(ref).astype(float), axis=0)
    mo = np.mean(np.stack(ours).astype(float), axis=0)
    assert np.abs(mr - mo).max() < 255 * 0.04, "per-pixel mean of imagined frames differs (all-skills goal)"


def test_cfg1_protocol_live(ray, oracle_lib):
    """BASELINE config 1 run live: the frozen cfg1 trace is what the reference produces today."""
    from tests import golden_util as gu
    d = gu.load("cfg1_21x21.npz")
    env = ray.CraftingWorldEnvRay(size=(21, 21), selected_tasks=['ChopTree', 'BuildHouse'], number_of_tasks=2)
    env.seed(0)
    env.reset()
    g, r, c, h, _, _ = ref_shim.read_back(env)
    assert np.array_equal(g, d["grid0"][0]) and (r, c, h) == (d["r0"][0], d["c0"][0], d["hold0"][0])
