"""Per-CTA timeline of the LAST fused launch of a host-driven closed loop (CW_LIB_PATH -> the -DCW_TIMING build): when do the CTAs
of a single cw_host_step enter, and when are their step phases (= status bytes) done?"""
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
    d = d[d[:, 0] > 0]
    t0 = d[:, 0].min()
    u = lambda x: (x - t0) / 1e3
    q = lambda x: "min %.2f p50 %.2f p90 %.2f max %.2f" % (u(x.min()), u(np.percentile(x, 50)), u(np.percentile(x, 90)), u(x.max()))
    print("CTAs %d | entry: %s | tiles landed: %s | step phase done: %s | exit: %s" % (len(d), q(d[:, 0]), q(d[:, 2]), q(d[:, 3]), q(d[:, 7])))
