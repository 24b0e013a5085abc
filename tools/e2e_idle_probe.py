"""Host-driven single steps with an idle gap between them (a slow policy): is the status latency of a step the kernel's own, or
the previous launch's tail?  CW_HOST_TRACE=1 prints launch -> first / all status bytes."""
import sys, time, os
sys.path.insert(0, ".")
import numpy as np
import gym_craftingworld_b200 as cw
N = 4096
gap_us = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
env = cw.HostCraftingWorldEnv(N, seed=0, return_frames=False)
env.reset()
env.load_state(t=np.random.RandomState(1).randint(0, 300, N))
acts = np.random.RandomState(0).randint(0, 6, (128, N)).astype(np.uint8)
for k in range(50):
    env.step(acts[k])
t_step = 0.0
for k in range(2000):
    t0 = time.perf_counter()
    env.step(acts[k % 128])
    t_step += time.perf_counter() - t0
    end = time.perf_counter() + gap_us * 1e-6
    while time.perf_counter() < end:
        pass
print(f"gap {gap_us:5.1f} us: step call {t_step / 2000 * 1e6:6.2f} us")
env.close()
