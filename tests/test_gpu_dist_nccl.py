"""GPU, world_size 2 over NCCL (skipped on a single-GPU box): the one collective of the path on hardware.  Each rank steps
its shard of the worlds on its own B200 through the fused kernel; every EVERY steps StatsReducer snapshots the device
accumulator on the step stream and all-reduces the copy on a side stream.  Every rank must see, for every reduction, exactly
the global statistics the C oracle has after the same step -- i.e. the snapshot is consistent (all counters from one step)
and the NCCL sum is right."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                                                  reason="needs >= 2 GPUs")]

TOTAL, K, EVERY, SEED, SIZE, MAX_STEPS = 6000, 96, 16, 77, 9, 7


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _actions():
    return np.random.RandomState(5).randint(0, 6, (K, TOTAL)).astype(np.uint8)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import gym_craftingworld_b200 as cw
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lo, n = cw.shard_range(TOTAL, rank, world)
        env = cw.BatchedCraftingWorldEnv(n, size=(SIZE, SIZE), max_steps=MAX_STEPS, seed=SEED, device=dev, env_id_base=lo)
        env.reset()
        red = cw.StatsReducer(env, every=EVERY)
        acts = torch.from_numpy(_actions()[:, lo:lo + n]).to(dev)
        totals = []
        for k in range(K):                                       # no host sync inside the loop: launches run ahead of the reductions
            env.step(acts[k])
            red.step()
            if (k + 1) % EVERY == 0:
                totals.append(red.total().clone())
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), totals=torch.stack(totals).cpu().numpy(), local=env.stats.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_stats_allreduce_matches_oracle(tmp_path):
    import torch.multiprocessing as mp
    from oracle import native
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    ob = native.OracleBatch(native.make_config(H=SIZE, W=SIZE, max_steps=MAX_STEPS), TOTAL, seed=SEED)
    ob.reset()
    acts = _actions()
    want = []
    for k in range(K):
        ob.step_full(acts[k], auto_reset=True)
        if (k + 1) % EVERY == 0:
            want.append(ob.stats.copy())
    want = np.stack(want)
    assert want[-1][0] > 1000                                      # thousands of episodes finished
    for r in range(world):
        assert np.array_equal(parts[r]["totals"], want), f"rank {r}: reduced statistics differ from the oracle's global ones"
    assert np.array_equal(parts[0]["local"] + parts[1]["local"], ob.stats)
