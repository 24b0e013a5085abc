"""End-to-end step latency of the host-buffer API (cw_host_step) at a given batch: delta transport, frames left on device."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import gym_craftingworld_b200 as cw

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
acts = np.random.RandomState(0).randint(0, 6, (128, N)).astype(np.uint8)
for variant in ("delta", "device"):
    env = cw.HostCraftingWorldEnv(N, size=(21, 21), seed=0, return_frames=variant != "device", transport="delta" if variant == "delta" else "frames")
    env.reset()
    for k in range(50):
        env.step(acts[k % 128])
    t0 = time.perf_counter()
    for k in range(steps):
        env.step(acts[k % 128])
    dt = time.perf_counter() - t0
    print(f"{variant:7s} N={N}: {dt / steps * 1e6:7.2f} us/step  {N * steps / dt / 1e6:7.1f} M env-steps/s", flush=True)
    env.close()
