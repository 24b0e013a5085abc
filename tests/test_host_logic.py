"""CPU: host-side logic of the facade -- constructor-argument handling mirrors the reference (ray.py:59-83),
sharding arithmetic, spaces metadata, and the no-CPU-fallback guarantee."""
import numpy as np
import pytest
import torch

import gym_craftingworld_b200 as cw
from gym_craftingworld_b200 import spaces
from gym_craftingworld_b200.env import make_config


def test_default_config_matches_reference_defaults():
    cfg = make_config()
    assert (cfg.H, cfg.W, cfg.cell_stride, cfg.max_steps) == (21, 21, 448, 300)
    assert (cfg.subset_reward, cfg.stacking, cfg.n_selected, cfg.number_of_tasks) == (0, 1, 9, 9)
    assert list(cfg.selected)[:9] == list(range(9))
    assert cw.TASK_LIST[4] == "ChopRock" and cw.TASK_LIST[8] == "MoveSticks"


def test_number_of_tasks_is_clipped_like_upstream():
    cfg = make_config(selected_tasks=["ChopTree", "BuildHouse"], number_of_tasks=5)     # ray.py:80-81
    assert cfg.number_of_tasks == 2 and list(cfg.selected)[:2] == [3, 2]
    assert make_config(reward_style="anything").subset_reward == 1                     # ray.py:71-74
    assert make_config(stacking=1).stacking == 0                                       # `is True` upstream, ray.py:169
    assert make_config(size=(32, 32)).cell_stride == 1024


def test_task_bits_follow_task_list_index():
    custom = list(reversed(cw.TASK_LIST))
    cfg = make_config(task_list=custom, selected_tasks=["MakeBread"])                  # ray.py:174: task_list.index(...)
    assert cfg.selected[0] == 8


@pytest.mark.parametrize("kw", [dict(size=(21, 20)), dict(size=(2, 2)), dict(size=(65, 65)), dict(max_steps=0),
                                dict(selected_tasks=["Fly"]), dict(selected_tasks=[]), dict(task_list=["a", "b"]),
                                dict(number_of_tasks=0)])
def test_bad_arguments_raise_value_error(kw):
    with pytest.raises(ValueError):
        make_config(**kw)


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not degrade to a CPU implementation."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        cw.BatchedCraftingWorldEnv(4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cw.HostCraftingWorldEnv(4)


def test_unsupported_side_channels_are_explicit():
    with pytest.raises(NotImplementedError):
        cw.BatchedCraftingWorldEnv(4, store_gif=True)
    with pytest.raises(ValueError):
        cw.BatchedCraftingWorldEnv(4, obs_mode="ascii")
    with pytest.raises(ValueError):
        cw.BatchedCraftingWorldEnv(4, render="lazy")
    with pytest.raises(ValueError):                                # incremental rendering patches ONE persistent pixel buffer
        cw.BatchedCraftingWorldEnv(4, render="incremental", obs_mode="compact")
    with pytest.raises(ValueError):
        cw.BatchedCraftingWorldEnv(4, render="incremental", obs_buffers=2)
    with pytest.raises(ValueError):
        cw.HostCraftingWorldEnv(4, transport="carrier-pigeon")


def test_product_never_imports_the_oracle():
    import os
    import re
    pkg = os.path.dirname(cw.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libcw_oracle" not in src, f


def test_shard_range_partitions_exactly():
    for total in (1, 7, 4096, 1 << 20, 1000003):
        for world in (1, 2, 3, 8):
            parts = [cw.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(n for _, n in parts) == total
            for (lo, n), (lo2, _) in zip(parts, parts[1:]):
                assert lo + n == lo2
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 1


def test_spaces_metadata():
    d = spaces.Discrete(6)
    assert d.n == 6 and all(0 <= d.sample() < 6 for _ in range(50)) and d.contains(5) and not d.contains(6)
    b = spaces.Box(0, 255, (84, 84, 3), np.uint8)
    assert b.shape == (84, 84, 3) and b.low == 0 and b.high == 255
    assert spaces.Dict({"a": b})["a"] is b


def test_gif_recorder_writes_episode_gifs(tmp_path):
    """Host-side GIF dump (the role of allow_gif_storage, ray.py:565-597, 769-782) on synthetic frames."""
    from PIL import Image
    rec = cw.GifRecorder(index=1, directory=str(tmp_path), env_id=7, scale=2)
    rng = np.random.RandomState(0)
    for t in range(6):
        frames = rng.randint(0, 255, (3, 20, 20, 3)).astype(np.uint8)
        done = np.array([False, t == 4, False])
        rec.capture({"observation": frames}, done)
    # auto-reset semantics: the frame returned WITH done is already the next episode's first frame
    assert len(rec.saved) == 1 and rec.saved[0].endswith("E0(3).gif") and len(rec.frames) == 2
    with Image.open(rec.saved[0]) as im:
        assert im.n_frames == 4 and im.size == (40, 40)
    assert cw.register_envs() == []                       # neither gym nor gymnasium is installed here


def test_registration_mirrors_the_references_three_ids():
    """`gym_craftingworld/__init__.py:5-18` registers craftingworld-v3 / craftingworldflat-v3 / craftingworldonehot-v3 with
    kwargs stacking=True, render_save_rate=10: the batched mirrors register the same three (tagged -b200, or under the
    reference's own ids), same kwargs + num_envs, and every entry point resolves to a class that accepts them."""
    import importlib
    import inspect
    seen = {}
    ids = cw.register_envs(num_envs=64, register=lambda id, entry_point, kwargs: seen.__setitem__(id, (entry_point, kwargs)))
    assert ids == ["craftingworld-b200-v3", "craftingworldflat-b200-v3", "craftingworldonehot-b200-v3"]
    seen_ref = {}
    assert cw.register_envs(reference_ids=True, register=lambda id, entry_point, kwargs: seen_ref.__setitem__(id, (entry_point, kwargs))) == \
        ["craftingworld-v3", "craftingworldflat-v3", "craftingworldonehot-v3"]
    for env_id, (entry, kwargs) in seen.items():
        assert kwargs == {"num_envs": 64, "stacking": True, "render_save_rate": 10}
        mod, cls = entry.split(":")
        klass = getattr(importlib.import_module(mod), cls)
        params = inspect.signature(klass.__init__).parameters
        assert "num_envs" in params and (("stacking" in params and "render_save_rate" in params) or "kw" in params or "args" in params)
    from oracle import ref_shim
    if ref_shim.reference_available():                    # the reference's own registry: same ids, same kwargs (minus num_envs)
        ref_shim.load_reference()
        import sys
        reg = sys.modules["gym.envs.registration"].registry
        for ref_id, (ref_cls, _) in cw.vector.REGISTRATIONS.items():
            entry, kwargs = reg[ref_id]
            assert entry.endswith(":" + ref_cls) and kwargs == {k: v for k, v in seen_ref[ref_id][1].items() if k != "num_envs"}


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (CPU only): exactly one JSON line on stdout with the contract's keys."""
    import json
    import subprocess
    import sys
    root = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "3", "--warmup", "3"], cwd=root,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    import os
    installed = os.path.isfile(os.path.join(root, "oracle", "_ref", "gym_craftingworld", "envs", "craftingworld_ray.py"))
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == ("reference" if installed else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # both arms print the SAME config dict: a pure function of the command line
    import argparse
    sys.path.insert(0, root)
    import bench
    assert d["config"] == bench.workload_config(argparse.Namespace(workload="cfg2", envs=0, ring=0), 1) and "method" in d


def test_tools_and_entry_points_compile():
    """bench.py, __graft_entry__.py and every probe under tools/ at least byte-compile (they only run on the GPU box)."""
    import glob
    import os
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")] + sorted(glob.glob(os.path.join(root, "tools", "*.py")))
    assert len(files) >= 8
    for f in files:
        py_compile.compile(f, doraise=True)


def test_numa_binding_helper_parses_topology(tmp_path, monkeypatch):
    """bind_to_gpu_numa_node: cpulist parsing, and a no-op (None) whenever the topology is unknown -- as here, without a GPU."""
    from gym_craftingworld_b200 import dist
    assert dist._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert dist._parse_cpulist("") == set()
    assert dist.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
