"""Why is bench.py's e2e leg slower than tools/e2e_probe.py on the same box?  Run the same closed loop after optional preambles."""
import sys, time, os
import numpy as np
sys.path.insert(0, ".")
import torch
import gym_craftingworld_b200 as cw

N = 4096
acts = np.random.RandomState(0).randint(0, 6, (128, N)).astype(np.uint8)

def loop(tag, steps=1000):
    env = cw.HostCraftingWorldEnv(N, size=(21, 21), seed=0, return_frames=False)
    env.reset()
    env.load_state(t=np.random.RandomState(1).randint(0, 300, N))
    for k in range(20):
        env.step(acts[k])
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for k in range(steps):
            env.step(acts[k % 128])
        env.sync()
        best = min(best, time.perf_counter() - t0)
    print(f"{tag:40s}: {best / steps * 1e6:6.2f} us/step", flush=True)
    env.close()

loop("fresh process")
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
if mode in ("all", "graph"):
    benv = cw.BatchedCraftingWorldEnv(N, seed=0, obs_buffers=4)
    benv.reset()
    tape = torch.randint(0, 6, (128, N), device="cuda", dtype=torch.uint8)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        benv.step(tape[0], chain_pos=0)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for k in range(128):
                benv.step(tape[k], chain_pos=k)
        for _ in range(20):
            g.replay()
        s.synchronize()
    loop("after a 128-step graph (graph alive)")
    del g, benv
    torch.cuda.empty_cache()
    loop("after deleting the graph")
if mode in ("all", "big"):
    benv = cw.BatchedCraftingWorldEnv(131072, seed=0, obs_buffers=2)
    benv.reset()
    benv.step(torch.zeros(131072, dtype=torch.uint8, device="cuda"))
    torch.cuda.synchronize()
    del benv
    torch.cuda.empty_cache()
    loop("after a 131072-world env (5.5 GB of frames)")
import gc
gc.disable()
loop("gc disabled")
