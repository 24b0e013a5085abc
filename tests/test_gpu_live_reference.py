"""GPU vs the LIVE, UNMODIFIED reference (`CraftingWorldEnvRay`), side by side on the same box: the reference package travels
to the GPU box as `oracle/_ref` (installed by `oracle/build_ref.py`; git-ignored).  BASELINE.json's parity protocol: the
reference generates the states, they are uploaded to the device, both sides get identical action sequences; grids, positions,
inventory, achieved vectors, rewards, dones and every rendered pixel must match at every step."""
import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_shim.reference_available(), reason="reference package not installed (oracle/_ref)")]


@pytest.fixture(scope="module")
def ray():
    return ref_shim.load_reference()


@pytest.fixture(scope="module")
def cw():
    import gym_craftingworld_b200 as m
    return m


def _upload(cw, envs, size, max_steps, reward_style=None, **kw):
    states = [ref_shim.read_back(e) for e in envs]
    desired = [ref_shim.bits_to_mask(e.desired_goal_vector[0]) for e in envs]
    dev = cw.BatchedCraftingWorldEnv(len(envs), size=(size, size), max_steps=max_steps, auto_reset=False, seed=0,
                                     reward_style=reward_style, **kw)
    dev.load_state(np.stack([s[0] for s in states]), [s[1] for s in states], [s[2] for s in states], [s[3] for s in states], desired)
    return dev


@pytest.mark.parametrize("size,max_steps,subset", [(21, 300, False), (8, 25, True)])
def test_kernels_step_beside_the_live_reference(ray, cw, size, max_steps, subset):
    """Worlds from the reference's own reset() (seed(i), as BASELINE config 2 prescribes), stepped past `done`."""
    B, T = 24, 160 if size == 21 else 90
    style = "s" if subset else None
    envs = []
    for i in range(B):
        e = ray.CraftingWorldEnvRay(size=(size, size), max_steps=max_steps, reward_style=style)
        e.seed(i)
        e.reset()
        envs.append(e)
    for mode in ("fused", "incremental"):
        for i, e in enumerate(envs):                               # the same start for both device render paths
            e.seed(i); e.reset()
        dev = _upload(cw, envs, size, max_steps, style, render="incremental" if mode == "incremental" else "full")
        assert np.array_equal(dev.obs.cpu().numpy(), np.stack([e.obs_image for e in envs]).astype(np.uint8))
        rng = np.random.RandomState(7)
        for t in range(T):
            a = rng.randint(0, 6, B)
            obs, reward, done, info = dev.step(torch.from_numpy(a).cuda())
            st = dev.export_state()
            frames = obs["observation"].cpu().numpy()
            ach = info["achieved_goal"].cpu().numpy()
            for b, e in enumerate(envs):
                o, rw, dn, inf = e.step(int(a[b]))
                g, r, c, h, amask, px = ref_shim.read_back(e)
                where = f"{mode} world {b} step {t}"
                assert np.array_equal(st["grid"][b], g), where
                assert (st["r"][b], st["c"][b], st["hold"][b]) == (r, c, h), where
                assert st["achieved"][b] == amask and np.array_equal(ach[b], np.asarray(inf["achieved_goal"]).reshape(-1)), where
                assert int(reward[b]) == rw and bool(done[b]) == dn, where
                assert np.array_equal(frames[b], px), where
                assert int(st["t"][b]) == e.step_num, where


def test_host_buffer_api_beside_the_live_reference(ray, cw):
    """The same protocol through the host-buffer C entry points (cw_host_load_state / cw_host_step): delta transport keeps the
    host frames equal to the reference's obs_image at every step."""
    B, size, T = 16, 21, 120
    envs = []
    for i in range(B):
        e = ray.CraftingWorldEnvRay(size=(size, size))
        e.seed(100 + i)
        e.reset()
        envs.append(e)
    states = [ref_shim.read_back(e) for e in envs]
    henv = cw.HostCraftingWorldEnv(B, size=(size, size), auto_reset=False, transport="delta")
    henv.reset()
    agent = np.array([s[1] | (s[2] << 8) | (s[3] << 16) for s in states], np.uint32)
    goal = np.array([ref_shim.bits_to_mask(e.desired_goal_vector[0]) << 16 for e in envs], np.uint32)
    obs = henv.load_state(np.stack([s[0] for s in states]), agent, goal, np.zeros(B, np.int32))
    assert np.array_equal(obs["observation"], np.stack([e.obs_image for e in envs]).astype(np.uint8))
    rng = np.random.RandomState(8)
    for t in range(T):
        a = rng.randint(0, 6, B)
        obs, reward, done, _ = henv.step(a)
        for b, e in enumerate(envs):
            _, rw, dn, _ = e.step(int(a[b]))
            assert int(reward[b]) == rw and bool(done[b]) == dn, (b, t)
            assert np.array_equal(obs["observation"][b], e.obs_image.astype(np.uint8)), (b, t)
    henv.close()


def test_pipelined_host_step_beside_the_live_reference(ray, cw):
    """Device-consumer transport (two-launch pipeline: step launch + render launch of its snapshot) beside the live
    reference: reward / done after every call, and the device frames -- fetched -- against the reference's obs_image."""
    B, size, T = 40, 21, 150
    envs = []
    for i in range(B):
        e = ray.CraftingWorldEnvRay(size=(size, size))
        e.seed(300 + i)
        e.reset()
        envs.append(e)
    states = [ref_shim.read_back(e) for e in envs]
    henv = cw.HostCraftingWorldEnv(B, size=(size, size), auto_reset=False, return_frames=False)
    henv.reset()
    agent = np.array([s[1] | (s[2] << 8) | (s[3] << 16) for s in states], np.uint32)
    goal = np.array([ref_shim.bits_to_mask(e.desired_goal_vector[0]) << 16 for e in envs], np.uint32)
    henv.load_state(np.stack([s[0] for s in states]), agent, goal, np.zeros(B, np.int32))
    assert np.array_equal(henv.fetch_frames()[0], np.stack([e.obs_image for e in envs]).astype(np.uint8))
    rng = np.random.RandomState(9)
    for t in range(T):
        a = rng.randint(0, 6, B)
        _, reward, done, _ = henv.step(a)
        for b, e in enumerate(envs):
            _, rw, dn, _ = e.step(int(a[b]))
            assert int(reward[b]) == rw and bool(done[b]) == dn, (b, t)
        if t % 5 == 4 or t == T - 1:
            frames = henv.fetch_frames()[0]
            for b, e in enumerate(envs):
                assert np.array_equal(frames[b], e.obs_image.astype(np.uint8)), (b, t)
    henv.close()
