"""Experiment: how fast can ANY kernel write an 86.7 MB frame buffer per launch in a 128-step CUDA graph on this GPU?
(torch fill_ = plain vectorised STG stream; ring of 4 buffers like bench.py cfg2)."""
import sys, torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
fb = 21168
bufs = [torch.empty(N * fb, dtype=torch.uint8, device="cuda") for _ in range(4)]
src = torch.empty(N * fb, dtype=torch.uint8, device="cuda")
def bench(fn, name, steps=2560):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for k in range(8): fn(k)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for k in range(128): fn(k)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(steps // 128): g.replay()
        e1.record(s); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / steps
    print(f"{name}: {us:.2f} us per launch, {N*fb/us/1e3:.0f} GB/s written")
bench(lambda k: bufs[k % 4].fill_(7), "fill_ (write only)")
bench(lambda k: bufs[k % 4].copy_(src), "copy_ (read+write)")
