"""Where does a K-step window go?  (CW_LIB_PATH -> the -DCW_TIMING build.)  One CUDA graph of K = 20 chained fused launches at config 2,
replayed from an idle GPU: per position first / median entry, median 'step phase done', first and last exit, relative to
position 0's first entry; and the CUDA-event time of the same replay."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import gym_craftingworld_b200 as cw
from gym_craftingworld_b200 import _lib
lib = _lib.load()
N, P, ROWS = int(os.environ.get("N", "4096")), int(os.environ.get("K", "20")), 1024
env = cw.BatchedCraftingWorldEnv(N, seed=0, obs_buffers=4)
env.reset()
env.t.copy_(torch.randint(0, 300, (N,), device="cuda", dtype=torch.int32))
tape = torch.randint(0, 6, (P, N), device="cuda", dtype=torch.uint8)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    env.step(tape[0], chain_pos=0)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for k in range(P):
            env.step(tape[k], chain_pos=k)
    for _ in range(3):
        g.replay()
torch.cuda.synchronize()
dbg = torch.zeros((P * ROWS, 16), dtype=torch.int64, device="cuda")
lib.cw_debug_set_timing.argtypes = [C.c_void_p]
assert lib.cw_debug_set_timing_rows_per_position(ROWS) == 0 and lib.cw_debug_set_timing(dbg.data_ptr()) == 0
for rep in range(3):
    dbg.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        torch.cuda._sleep(600000)
        e0.record(s)
        g.replay()
        e1.record(s)
    torch.cuda.synchronize()
    d = dbg.cpu().numpy().astype(np.float64).reshape(P, ROWS, 16)
    t0 = d[0][d[0][:, 0] > 0][:, 0].min()
    u = lambda x: (x - t0) / 1e3
    print(f"replay {rep}: CUDA events {e0.elapsed_time(e1) * 1e3:.1f} us for {P} steps")
    if rep < 2:
        continue
    print("pos | CTAs | entry first / median / last | step phase done median | exit first / last | last exit - previous")
    prev = None
    for p in range(P):
        r = d[p]
        r = r[r[:, 0] > 0]
        le = u(r[:, 7].max())
        print("%3d | %4d | %7.2f / %7.2f / %7.2f | %7.2f | %7.2f / %7.2f | %s" % (p, len(r), u(r[:, 0].min()), u(np.median(r[:, 0])), u(r[:, 0].max()),
              u(np.median(r[:, 3])), u(r[:, 7].min()), le, "%.2f" % (le - prev) if prev is not None else "-"))
        prev = le
