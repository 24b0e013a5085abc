"""Phase times of the device-consumer single step (CW_HOST_TRACE=1): pipelined (two launches, two streams) vs fused chained launch."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
os.environ["CW_HOST_TRACE"] = "1"
import gym_craftingworld_b200 as cw

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
work_us = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0       # host "policy" time between two steps
acts = np.random.RandomState(0).randint(0, 6, (128, N)).astype(np.uint8)
env = cw.HostCraftingWorldEnv(N, size=(21, 21), seed=0, return_frames=False)
env.reset()
env.load_state(t=np.random.RandomState(1).randint(0, 300, N))
for k in range(100):
    env.step(acts[k % 128])
env.sync()
t0 = time.perf_counter()
for k in range(steps):
    env.step(acts[k % 128])
    if work_us:
        t1 = time.perf_counter()
        while (time.perf_counter() - t1) * 1e6 < work_us:
            pass
env.sync()
dt = time.perf_counter() - t0
print(f"pipe={os.environ.get('CW_HOST_PIPE', '1')} N={N} work={work_us} us: {dt / steps * 1e6:7.2f} us/step  {N * steps / dt / 1e6:7.1f} M env-steps/s", flush=True)
env.close()
