"""Soak test of the two-launch pipeline of the host-driven step (and of cw_step_chained): tens of thousands of steps against the
same steps taken without it -- reward / done after EVERY call, frames / goal frames / state at intervals, statistics at the end --
for tiny batches (every CTA of several launches co-resident), 1-step episodes (every world re-seeded in every step), the bench
shape and a multi-wave batch."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import gym_craftingworld_b200 as cw


def host_pair(N, size, max_steps):
    os.environ["CW_HOST_PIPE"] = "1"
    a = cw.HostCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, seed=5, return_frames=False)
    os.environ["CW_HOST_PIPE"] = "0"
    b = cw.HostCraftingWorldEnv(N, size=(size, size), max_steps=max_steps, seed=5, return_frames=False)
    os.environ["CW_HOST_PIPE"] = "1"
    return a, b


for N, size, max_steps, steps in ((33, 5, 4, 30000), (700, 21, 1, 8000), (4096, 21, 300, 20000), (4096, 21, 7, 8000), (12000, 21, 40, 3000), (512, 32, 25, 5000)):
    a, b = host_pair(N, size, max_steps)
    a.reset(); b.reset()
    rng = np.random.RandomState(N)
    bad = 0
    t0 = time.time()
    for k in range(steps):
        act = rng.randint(0, 6, N).astype(np.uint8)
        _, ra, da, _ = a.step(act)
        ra, da = ra.copy(), da.copy()
        _, rb, db, _ = b.step(act)
        bad += int(not (np.array_equal(ra, rb) and np.array_equal(da, db)))
        if k % 499 == 498 or k == steps - 1:
            fa, ga = a.fetch_frames()
            fb, gb = b.fetch_frames()
            bad += int(not np.array_equal(fa, fb)) + int(not np.array_equal(ga, gb))
        if k % 1500 == 1499:                                       # leave the pipe for an open-loop run and come back
            tape = rng.randint(0, 6, (9, N)).astype(np.uint8)
            _, r1, d1, _ = a.step_many(tape)
            r1, d1 = r1.copy(), d1.copy()
            _, r2, d2, _ = b.step_many(tape)
            bad += int(not (np.array_equal(r1, r2) and np.array_equal(d1, d2)))
    bad += int(not np.array_equal(a.stats(), b.stats()))
    print(f"host pipeline N={N} {size}x{size} max_steps={max_steps}: {steps} steps, episodes {int(a.stats()[0])}, mismatches {bad} ({time.time() - t0:.0f} s)", flush=True)
    a.close(); b.close()
    assert bad == 0

for N, size, max_steps, reps in ((96, 5, 3, 300), (65536, 21, 300, 60), (5000, 21, 2, 100), (20000, 9, 12, 100)):
    K = 128
    tape = torch.randint(0, 6, (K, N), device="cuda", dtype=torch.uint8)
    kw = dict(size=(size, size), max_steps=max_steps, seed=3, obs_mode="compact")
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
                for key in ("grid", "init_grid", "agent", "goal", "t", "episode", "reward", "done", "stats_raw"):
                    bad += int(not torch.equal(getattr(a, key), getattr(b, key)))
    print(f"cw_step_chained N={N} {size}x{size} max_steps={max_steps}: {reps * K} chained steps, mismatches {bad}", flush=True)
    assert bad == 0
print("soak ok")
