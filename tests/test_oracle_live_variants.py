"""CPU, builder container only: the reference's OTHER registered env classes -- CraftingWorldEnvOneHot
(carftingworld_onehot.py) and CraftingWorldEnvFlat (craftingworld_flat.py) -- run live against the oracle restatement,
and against the fixtures frozen from them (tests/golden/variants_*.npz, which travel to the GPU box).  SURVEY 8(f)-2."""
import os

import numpy as np
import pytest

from oracle import compact, ref_shim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
live = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference source tree not present")


def drive(env_cls, kw, seed, actions, flat):
    """Run a reference variant env; return per-step compact states read back from it, its observations, rewards, dones."""
    env = env_cls(**kw)
    env.seed(seed)
    obs0 = env.reset()
    g, r, c, h = ref_shim.onehot_to_compact(env.obs_one_hot)
    out = dict(grid0=g, r0=r, c0=c, hold0=h, desired=ref_shim.bits_to_mask(env.desired_goal_vector[0]),
               obs0=np.asarray(obs0 if flat else obs0["observation"]).copy(), grid=[], r=[], c=[], hold=[], achieved=[],
               reward=[], done=[], obs=[])
    for a in actions:
        o, rw, dn, _ = env.step(int(a))
        g, r, c, h = ref_shim.onehot_to_compact(env.obs_one_hot)
        out["grid"].append(g); out["r"].append(r); out["c"].append(c); out["hold"].append(h)
        out["achieved"].append(ref_shim.bits_to_mask(env.achieved_goal_vector[0]))
        out["reward"].append(rw); out["done"].append(dn)
        out["obs"].append(np.asarray(o if flat else o["observation"]).copy())
    return {k: np.asarray(v) for k, v in out.items()}


def check_against_oracle(d, size, max_steps, flat):
    """the compact restatement reproduces the variant's trajectory; its observation is the one-hot state / the RGB frame"""
    pc = compact.Config(H=size, W=size, max_steps=max_steps)
    st = compact.EnvState(d["grid0"].copy(), d["grid0"].copy(), int(d["r0"]), int(d["c0"]), int(d["hold0"]), 0, int(d["desired"]))
    first = compact.render(st.grid, st.r, st.c, st.hold) if flat else ref_shim.compact_to_onehot(st.grid, st.r, st.c, st.hold)
    assert np.array_equal(np.asarray(d["obs0"]), first)
    for t, a in enumerate(d["actions"]):
        rw, dn, _ = compact.step_env(st, int(a), pc)
        assert (rw, bool(dn)) == (int(d["reward"][t]), bool(d["done"][t])), t
        assert np.array_equal(st.grid, d["grid"][t]) and (st.r, st.c, st.hold) == (d["r"][t], d["c"][t], d["hold"][t]), t
        assert st.achieved == d["achieved"][t], t
        want = compact.render(st.grid, st.r, st.c, st.hold) if flat else ref_shim.compact_to_onehot(st.grid, st.r, st.c, st.hold)
        assert np.array_equal(np.asarray(d["obs"][t]), want), t


CASES = [("onehot", 6, 30, 11, 90), ("onehot", 21, 300, 12, 200), ("flat", 8, 100, 13, 150), ("flat", 5, 12, 14, 60)]


def variant_class(kind):
    ref_shim.load_reference()
    if kind == "onehot":
        import gym_craftingworld.envs.carftingworld_onehot as m
        return m.CraftingWorldEnvOneHot
    import gym_craftingworld.envs.craftingworld_flat as m
    return m.CraftingWorldEnvFlat


@live
@pytest.mark.parametrize("kind,size,max_steps,seed,T", CASES)
def test_variant_envs_live(kind, size, max_steps, seed, T):
    actions = np.random.RandomState(seed).randint(0, 6, T)
    d = drive(variant_class(kind), dict(size=(size, size), max_steps=max_steps), seed, actions, flat=kind == "flat")
    d["actions"] = actions
    check_against_oracle(d, size, max_steps, flat=kind == "flat")


@live
def test_flat_defaults_live():
    env = variant_class("flat")()
    assert (env.STATE_W, env.STATE_H, env.MAX_STEPS) == (8, 8, 100)            # craftingworld_flat.py:40-43
    assert np.asarray(env.reset()).shape == (32, 32, 3)


@pytest.mark.parametrize("name", ["variants_onehot_6x6.npz", "variants_flat_8x8.npz"])
def test_variant_fixtures_match_oracle(name):
    """the frozen copies of two of the runs above (made by tests/golden/make_golden_variants.py)"""
    d = dict(np.load(os.path.join(GOLDEN, name)))
    check_against_oracle(d, int(d["size"]), int(d["max_steps"]), flat=bool(d["flat"]))
