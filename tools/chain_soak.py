"""Soak test of chained launches: many graph replays, chained vs independent launches must stay bit-identical
(state, statistics, every frame buffer), for a tiny batch (several chain positions co-resident), the bench shape and a
persistent launch."""
import sys
import torch
sys.path.insert(0, ".")
import gym_craftingworld_b200 as cw

for N, size, max_steps, ring, reps in ((64, 7, 5, 1, 300), (64, 7, 5, 3, 300), (4096, 21, 300, 4, 100), (2048, 21, 20, 2, 100), (40000, 21, 30, 2, 20)):
    K = 128
    tape = torch.randint(0, 6, (K, N), device="cuda", dtype=torch.uint8)
    kw = dict(size=(size, size), max_steps=max_steps, seed=3, obs_buffers=ring)
    a, b = cw.BatchedCraftingWorldEnv(N, **kw), cw.BatchedCraftingWorldEnv(N, **kw)
    a.reset(); b.reset()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        a.step(tape[0], chain_pos=0); b.step(tape[0])
        ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga, stream=s):
            for k in range(K):
                a.step(tape[k], chain_pos=k)
        with torch.cuda.graph(gb, stream=s):
            for k in range(K):
                b.step(tape[k])
        bad = 0
        for r in range(reps):
            ga.replay(); gb.replay()
            if r % 10 == 9 or r == reps - 1:
                s.synchronize()
                for key in ("grid", "init_grid", "agent", "goal", "t", "episode", "reward", "desired_goal", "init_obs", "stats_raw"):
                    bad += int(not torch.equal(getattr(a, key), getattr(b, key)))
                for i in range(ring):
                    bad += int(not torch.equal(a._obs_ring[i], b._obs_ring[i]))
    print(f"N={N} {size}x{size} max_steps={max_steps} ring={ring}: {reps * K} chained steps, mismatches {bad}", flush=True)
    assert bad == 0
print("soak ok")
