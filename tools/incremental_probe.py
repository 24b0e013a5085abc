"""env-steps/s with pixel observations kept current by incremental rendering (cw_step_render_edit), CUDA-graph replay."""
import sys
import torch
sys.path.insert(0, ".")
import gym_craftingworld_b200 as cw

cases = [(4096, 21, True), (65536, 21, True), (65536, 21, False), (131072, 21, True), (16384, 32, True)]
for N, size, ar in cases:
    env = cw.BatchedCraftingWorldEnv(N, size=(size, size), seed=0, render="incremental", auto_reset=ar)
    env.reset()
    tape = torch.randint(0, 6, (128, N), device="cuda", dtype=torch.uint8)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for k in range(8):
            env.step(tape[k])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for k in range(128):
                env.step(tape[k])
        g.replay(); g.replay(); g.replay()                       # past the first synchronised time-out
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / (reps * 128) * 1e-3
    print(f"incremental auto_reset={ar} N={N:6d} {size}x{size}: {t*1e6:7.2f} us/step  {N/t/1e9:6.2f} G env-steps/s with current pixel frames in HBM", flush=True)
    del env
