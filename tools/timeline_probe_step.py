"""Experiment (CW_LIB_PATH -> a -DCW_TIMING build): per-warp timeline of cw_step_kernel at config 3."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import gym_craftingworld_b200 as cw
from gym_craftingworld_b200 import _lib
lib = _lib.load()
N = 65536
env = cw.BatchedCraftingWorldEnv(N, seed=0, obs_mode="compact", max_steps=int(os.environ.get("MAXS", "300")))
env.reset()
env.t.copy_(torch.randint(0, env.MAX_STEPS, (N,), device="cuda", dtype=torch.int32))      # staggered episodes: ~N/max_steps re-seeds per step
tape = torch.randint(0, 6, (128, N), device="cuda", dtype=torch.uint8)
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300):
    env.step(tape[k % 128])
dbg = torch.zeros((4096, 16), dtype=torch.int64, device="cuda")
lib.cw_debug_set_timing.argtypes = [C.c_void_p]
assert lib.cw_debug_set_timing(dbg.data_ptr()) == 0
torch.cuda.synchronize()
for rep in range(8):
    dbg.zero_()
    for k in range(4): env.step(tape[k])
    torch.cuda.synchronize()
    d = dbg.cpu().numpy()[:N // 32]                      # stepping warps only (refill CTAs follow)
    d = d[d[:, 0] > 0].astype(np.float64)
    base = d[:, 1].min()
    r = (d[:, :5] - base) / 1e3
    nres = d[:, 9].astype(int)
    hr = nres > 0
    f = lambda x: "%.2f" % x
    print("warps", len(d), "| start spread", f(r[:, 1].max()), "| step done mean", f(r[:, 2].mean()), "max", f(r[:, 2].max()),
          "| end mean", f(r[:, 4].mean()), "max", f(r[:, 4].max()), "| reset warps", int(hr.sum()), "resets", int(nres.sum()),
          "| reset duration mean", f((r[hr, 3] - r[hr, 2]).mean() if hr.any() else 0), "max", f((r[hr, 3] - r[hr, 2]).max() if hr.any() else 0),
          "| end max (reset warps)", f(r[hr, 4].max() if hr.any() else 0), "(others)", f(r[~hr, 4].max()))
