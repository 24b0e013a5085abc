"""End-to-end step latency of the host-buffer API (cw_host_step) at a given batch: delta transport (frames current in host
memory), device consumer (frames left in HBM), and cw_host_step_many (K steps per call)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import gym_craftingworld_b200 as cw

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
import os
SIZE = int(os.environ.get("SIZE", "21"))
stagger = (int(sys.argv[3]) if len(sys.argv) > 3 else 1) != 0
acts = np.random.RandomState(0).randint(0, 6, (128, N)).astype(np.uint8)
def make(variant):
    env = cw.HostCraftingWorldEnv(N, size=(SIZE, SIZE), seed=0, return_frames=variant != "device", transport="delta" if variant == "delta" else "frames")
    env.reset()
    if stagger:
        env.load_state(t=np.random.RandomState(1).randint(0, 300, N))     # staggered episodes: a steady stream of re-seeds
    return env


for variant in ("delta", "device"):
    env = make(variant)
    for k in range(50):
        env.step(acts[k % 128])
    t0 = time.perf_counter()
    for k in range(steps):
        env.step(acts[k % 128])
    env.sync()
    dt = time.perf_counter() - t0
    print(f"{variant:7s} N={N}: {dt / steps * 1e6:7.2f} us/step  {N * steps / dt / 1e6:7.1f} M env-steps/s", flush=True)
    env.close()
    env = make(variant)
    for K in (16, 128):
        reps = max(1, steps // K)
        env.step_many(acts[:K])
        t0 = time.perf_counter()
        for _ in range(reps):
            env.step_many(acts[:K])
        env.sync()
        dt = time.perf_counter() - t0
        print(f"{variant:7s} N={N} step_many K={K}: {dt / (reps * K) * 1e6:7.2f} us/step  {N * reps * K / dt / 1e6:7.1f} M env-steps/s", flush=True)
    env.close()
