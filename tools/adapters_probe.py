"""Roofline of the observation-adapter kernels: cw_onehot (12 B per cell), cw_render_alt (int16 3x3
sub-pixel frames) and cw_render (render only), achieved HBM GB/s against MEASURED_PEAKS.json."""
import json, os, sys
import torch
sys.path.insert(0, ".")
import gym_craftingworld_b200 as cw

peak = 6552.3
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak)

def timeit(fn, iters, per_graph=10):
    """device time per call, replayed from a CUDA graph of `per_graph` calls (the eager loop is CPU-bound at 4096 worlds)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(per_graph):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(max(1, iters // per_graph)):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (max(1, iters // per_graph) * per_graph) * 1e-3

for N in (4096, 65536):
    H = W = 21
    env = cw.BatchedCraftingWorldEnvAltObs(N, size=(W, H), seed=0)
    env.reset()
    iters = 200 if N <= 4096 else 40
    # buffers larger than L2 are written at 65536 worlds; at 4096 the outputs (21-25 MB) fit L2 -- noted in the output
    t = timeit(lambda: env.onehot(), iters)
    b = N * H * W * 12 + N * (H * W + 4)
    print(f"cw_onehot      N={N:6d}: {t*1e6:8.1f} us  {b/t/1e9:7.1f} GB/s  frac {b/t/1e9/peak:.3f}  (out {N*H*W*12/1e6:.0f} MB)")
    t = timeit(lambda: env.render_alt(env.grid, env.agent), iters)
    b = N * (3 * H + 3) * 3 * W * 3 * 2 + N * (H * W + 4)
    print(f"cw_render_alt  N={N:6d}: {t*1e6:8.1f} us  {b/t/1e9:7.1f} GB/s  frac {b/t/1e9/peak:.3f}  (out {N*(3*H+3)*3*W*6/1e6:.0f} MB)")
    env2 = cw.BatchedCraftingWorldEnv(N, size=(W, H), seed=0)
    env2.reset()
    t = timeit(lambda: env2.render(), iters)
    b = N * (48 * H * W + H * W + 4)
    print(f"cw_render      N={N:6d}: {t*1e6:8.1f} us  {b/t/1e9:7.1f} GB/s  frac {b/t/1e9/peak:.3f}  (out {N*48*H*W/1e6:.0f} MB)")
