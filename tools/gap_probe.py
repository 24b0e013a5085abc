"""Experiment (CW_LIB_PATH -> the -DCW_TIMING build, tools/build_timing.sh): what happens BETWEEN two dependent launches of the
compact step kernel?  Launch A steps 65536 worlds, launch B (next in the stream / graph, PDL) 32768 worlds of another env: B's
warps overwrite the first half of the stamp rows, A's second half survives, so one buffer shows A's end and B's start."""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np, torch
import gym_craftingworld_b200 as cw
from gym_craftingworld_b200 import _lib
lib = _lib.load()
import os
NA, NB = 65536, 32768
VAR = os.environ.get("VAR", "")
mk = lambda n: cw.BatchedCraftingWorldEnv(n, seed=0, obs_mode="compact", collect_stats="nostats" not in VAR, auto_reset="noreset" not in VAR)
A, B = mk(NA), mk(NB)
for e in (A, B):
    if "norecords" in VAR:
        e.reset_rec = e.reset_list = None
        e._refresh_state_struct()
    e.reset()
    e.t.copy_(torch.randint(0, 300, (e.num_envs,), device="cuda", dtype=torch.int32))
ta = torch.randint(0, 6, (8, NA), device="cuda", dtype=torch.uint8)
tb = torch.randint(0, 6, (8, NB), device="cuda", dtype=torch.uint8)
for k in range(8):
    A.step(ta[k]); B.step(tb[k])
dbg = torch.zeros((8192, 16), dtype=torch.int64, device="cuda")
lib.cw_debug_set_timing.argtypes = [C.c_void_p]
assert lib.cw_debug_set_timing(dbg.data_ptr()) == 0
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        A.step(ta[0]); A.step(ta[1]); B.step(tb[0])
print("variant:", VAR or "default", "CW_PDL=" + os.environ.get("CW_PDL", "1"))
for mode in ("graph",):
    for rep in range(4):
        dbg.zero_(); torch.cuda.synchronize()
        with torch.cuda.stream(s):
            if mode == "graph":
                g.replay()
            else:
                A.step(ta[0]); A.step(ta[1]); B.step(tb[0])
        torch.cuda.synchronize()
        d = dbg.cpu().numpy().astype(np.float64)
        a = d[NB // 32:NA // 32]; b = d[:NB // 32]
        a = a[a[:, 0] > 0]; b = b[b[:, 0] > 0]
        t0 = a[:, 1].min()
        f = lambda x: "%6.2f" % ((x - t0) / 1e3)
        print(mode, "| A: first past-wait 0.00, last end", f(a[:, 4].max()), "| B: first CTA resident", f(b[:, 0].min()), "last resident", f(b[:, 0].max()),
              "first past-wait", f(b[:, 1].min()), "last past-wait", f(b[:, 1].max()), "last end", f(b[:, 4].max()),
              "|| gap A.end -> B.first past-wait: %.2f us" % ((b[:, 1].min() - a[:, 4].max()) / 1e3))
