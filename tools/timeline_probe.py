"""Experiment (needs ab/lib_timing.so built with -DCW_TIMING and CW_LIB_PATH pointing at it): per-CTA timeline of
cw_env_kernel at config 2 in steady state (episodes desynchronised)."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import gym_craftingworld_b200 as cw
from gym_craftingworld_b200 import _lib
lib = _lib.load()
N = 4096
env = cw.BatchedCraftingWorldEnv(N, seed=0, obs_buffers=int(os.environ.get("RING", "4")), max_steps=int(os.environ.get("MAXS", "300")))
env.reset()
tape = torch.randint(0, 6, (128, N), device="cuda", dtype=torch.uint8)
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3000):
    env.step(tape[k % 128])
dbg = torch.zeros((1024, 16), dtype=torch.int64, device="cuda")
lib.cw_debug_set_timing.argtypes = [C.c_void_p]
assert lib.cw_debug_set_timing(dbg.data_ptr()) == 0
torch.cuda.synchronize()
for rep in range(12):
    dbg.zero_()
    for k in range(4): env.step(tape[k])       # back-to-back launches; the last one's stamps survive
    torch.cuda.synchronize()
    d = dbg.cpu().numpy()
    d = d[d[:, 0] > 0].astype(np.float64)
    base = d[:, 1].min()                       # first CTA past the PDL wait = real start of this launch's work
    r = (d[:, :9] - base) / 1e3
    npend = d[:, 9].astype(int)
    hr = npend > 0
    def f(x): return "%.2f" % x
    print("tiles landed (mean after wait release) %.2f | step done %.2f |" % ((r[:, 2] - r[:, 1]).mean(), (r[:, 3] - r[:, 2]).mean()), end=" ")
    print("launch: ctas", len(d), "work span (first past-wait -> last exit)", f(r[:, 7].max()), "us | exit mean", f(r[:, 7].mean()),
          "| no-reset CTAs: C1done", f(r[~hr, 4].mean()), "exit max", f(r[~hr, 7].max()),
          "| reset CTAs", int(hr.sum()), "pending worlds", int(npend.sum()))
    for i in np.flatnonzero(hr)[:6]:
        print("     reset CTA: past-wait %s tiles %s step %s | resetwarp done %s (busy %s) | C1done %s resetwait-> %s C2done %s exit %s  pending=%d"
              % (f(r[i,1]), f(r[i,2]), f(r[i,3]), f(r[i,8]), f(r[i,8]-r[i,3]), f(r[i,4]), f(r[i,5]), f(r[i,6]), f(r[i,7]), npend[i]))
