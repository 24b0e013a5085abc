"""Timeline of a CHAIN of fused launches (CW_LIB_PATH -> the -DCW_TIMING build, tools/build_timing.sh): every CTA of every chain
position stamps %globaltimer at entry (0), after its tiles landed (2), after the step phase (3), after composing the frames of
untouched worlds (4) and at exit (7).  One CUDA graph of P chained steps at config 2 is replayed; the table shows, per position,
when its first CTA entered, when its step phases were done, and when its last CTA left -- position i+1 steps and composes while
position i is still storing, and the positions leave one steady-state period apart.  (ncu cannot show this: it serialises launches.)"""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import gym_craftingworld_b200 as cw
from gym_craftingworld_b200 import _lib
lib = _lib.load()
N, P, ROWS = int(os.environ.get("N", "4096")), 24, 1024
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
with torch.cuda.stream(s):
    g.replay()
torch.cuda.synchronize()
d = dbg.cpu().numpy().astype(np.float64).reshape(P, ROWS, 16)
t0 = None
print(f"chained graph, {N} worlds, {P} positions; times in us relative to position 8's first CTA entry")
print("pos | CTAs | first entry | step phase done: first / median / last | untouched worlds composed (median) | last CTA exit | exit - previous exit")
prev_exit = None
for p in range(8, P):
    r = d[p]
    r = r[r[:, 0] > 0]
    if t0 is None:
        t0 = r[:, 0].min()
    u = lambda x: (x - t0) / 1e3
    exit_last = u(r[:, 7].max())
    print("%3d | %4d | %8.2f | %8.2f / %8.2f / %8.2f | %8.2f | %8.2f | %s" % (
        p, len(r), u(r[:, 0].min()), u(r[:, 3].min()), u(np.median(r[:, 3])), u(r[:, 3].max()), u(np.median(r[:, 4])), exit_last,
        "%.2f" % (exit_last - prev_exit) if prev_exit is not None else "-"))
    prev_exit = exit_last
