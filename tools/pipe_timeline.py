"""Per-CTA timeline of the LAST pipelined host step (CW_LIB_PATH -> the -DCW_TIMING build): the step launch's warps (rows 600..)
against the render launch's CTAs, one clock (%globaltimer)."""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np, torch
import gym_craftingworld_b200 as cw
from gym_craftingworld_b200 import _lib
lib = _lib.load()
N = 4096
env = cw.HostCraftingWorldEnv(N, seed=0, return_frames=False)
env.reset()
env.load_state(t=np.random.RandomState(1).randint(0, 300, N))
acts = np.random.RandomState(0).randint(0, 6, (128, N)).astype(np.uint8)
for k in range(300):
    env.step(acts[k % 128])
dbg = torch.zeros((1024, 16), dtype=torch.int64, device="cuda")
lib.cw_debug_set_timing.argtypes = [C.c_void_p]
assert lib.cw_debug_set_timing(dbg.data_ptr()) == 0
for rep in range(6):
    for k in range(50):
        env.step(acts[k])
    env.sync()
    d = dbg.cpu().numpy().astype(np.float64)
    r, s = d[:600], d[600:]
    r, s = r[r[:, 0] > 0], s[s[:, 0] > 0]
    t0 = s[:, 0].min()
    u = lambda x: (x - t0) / 1e3
    q = lambda x: "%.2f / %.2f / %.2f" % (u(x.min()), u(np.percentile(x, 50)), u(x.max()))
    print("step warps %d (min / p50 / max us after the first warp's entry): entry %s | predecessor acquired %s | stepped %s | status issued %s | slot free %s | "
          "copied %s | published %s | exit %s" % (len(s), q(s[:, 0]), q(s[:, 1]), q(s[:, 2]), q(s[:, 3]), q(s[:, 4]), q(s[:, 5]), q(s[:, 6]), q(s[:, 7])))
    print("   render CTAs %d: entry %s | tiles landed %s | exit %s" % (len(r), q(r[:, 0]), q(r[:, 2]), q(r[:, 7])))
