"""CPU, world_size 2 over gloo: the multi-rank plumbing.  Worlds shard by global index with no data-path
collective; the only exchange is the SUM all-reduce of the 24 x int64 statistics vector (StatsReducer).
The per-rank env work is done by the C oracle here (no GPU in this container); what is under test is the host
logic: shard_range, global-id keyed streams (sharded == unsharded), and the periodic reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_craftingworld_b200.dist import StatsReducer, shard_range
from oracle import native

TOTAL, K, EVERY, SEED = 96, 64, 16, 2024


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _actions():
    return np.random.RandomState(3).randint(0, 6, (K, TOTAL)).astype(np.uint8)


def _run_slice(lo, n):
    cfg = native.make_config(H=6, W=6, max_steps=9)
    ob = native.OracleBatch(cfg, n, seed=SEED, env_id_base=lo)
    ob.reset()
    return ob


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, n = shard_range(TOTAL, rank, world)
        ob = _run_slice(lo, n)
        stats = torch.from_numpy(ob.stats)                      # shares memory with the oracle's accumulator
        red = StatsReducer(stats, every=EVERY)
        acts = _actions()
        snapshots = []
        for k in range(K):
            ob.step_full(acts[k, lo:lo + n], auto_reset=True)
            red.step()
            if (k + 1) % EVERY == 0:
                snapshots.append(red.wait().clone().numpy())
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), grid=ob.grid, agent=ob.agent, goal=ob.goal, stats=ob.stats,
                 snapshots=np.stack(snapshots), lo=lo, n=n)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharding_and_stats_allreduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    # single-process ground truth over all worlds, snapshotting the stats at the same steps
    whole = _run_slice(0, TOTAL)
    acts = _actions()
    want_snapshots = []
    for k in range(K):
        whole.step_full(acts[k], auto_reset=True)
        if (k + 1) % EVERY == 0:
            want_snapshots.append(whole.stats.copy())
    assert [int(p["lo"]) for p in parts] == [0, 48] and [int(p["n"]) for p in parts] == [48, 48]
    for key in ("grid", "agent", "goal"):
        assert np.array_equal(np.concatenate([p[key] for p in parts]), getattr(whole, key)), key
    assert np.array_equal(parts[0]["stats"] + parts[1]["stats"], whole.stats)
    assert whole.stats[0] > 0
    for r in range(world):                                        # every rank sees the global sums
        assert np.array_equal(parts[r]["snapshots"], np.stack(want_snapshots)), f"rank {r}"


def test_stats_reducer_without_process_group_is_identity():
    stats = torch.arange(24, dtype=torch.int64)
    red = StatsReducer(stats, every=2)
    red.step()
    assert int(red.global_stats.sum()) == 0
    red.step()
    assert torch.equal(red.wait(), stats)
