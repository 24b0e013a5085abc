"""Compact step (BASELINE config 3 shape): graph of K single-step launches, ordinary (cw_step) vs chained per warp (cw_step_chained)."""
import sys, os
sys.path.insert(0, ".")
import numpy as np, torch
import gym_craftingworld_b200 as cw

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 128
acts = torch.from_numpy(np.random.RandomState(0).randint(0, 6, (K, N)).astype(np.uint8)).cuda()
for chained in (False, True):
    env = cw.BatchedCraftingWorldEnv(N, seed=0, obs_mode="compact")
    env.reset()
    env.t.copy_(torch.from_numpy(np.random.RandomState(1).randint(0, 300, N).astype(np.int32)).cuda())   # staggered episode clocks
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        env.step(acts[0]); env.step(acts[0], chain_pos=0)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for k in range(K):
                env.step(acts[k], chain_pos=k if chained else None)
        for _ in range(5):
            g.replay()
        s.synchronize()
        for reps in (1, 20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(600000)
            e0.record(s)
            for _ in range(reps):
                g.replay()
            e1.record(s)
            s.synchronize()
            ms = e0.elapsed_time(e1)
            print(f"N={N} K={K} chained={chained} order_every={os.environ.get('CW_STEP_CHAIN_ORDER_EVERY', '1')} replays={reps}: {ms * 1e3 / (reps * K):6.2f} us per step launch, "
                  f"{N * reps * K / ms / 1e6:8.2f} G env-steps/s", flush=True)
