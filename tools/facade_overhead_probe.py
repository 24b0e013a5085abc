"""Host cost of one eager BatchedCraftingWorldEnv.step() call (the call a PyTorch RL loop makes): wall clock per call with the
GPU kept busy only by the env itself, pixel and compact observations."""
import sys, time
sys.path.insert(0, ".")
import torch
import gym_craftingworld_b200 as cw

for mode, N in (("pixels", 4096), ("pixels", 256), ("compact", 65536), ("compact", 256)):
    env = cw.BatchedCraftingWorldEnv(N, seed=0, obs_mode=mode, obs_buffers=4 if mode == "pixels" else 1)
    env.reset()
    a = torch.randint(0, 6, (N,), device="cuda", dtype=torch.uint8)
    for _ in range(200):
        env.step(a)
    torch.cuda.synchronize()
    for reps in (2000,):
        t0 = time.perf_counter()
        for _ in range(reps):
            env.step(a)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{mode:8s} N={N:6d}: {1e6 * (t1 - t0) / reps:6.2f} us per step() call on the host, {1e6 * (t2 - t0) / reps:6.2f} us per step incl. the drain", flush=True)
