"""Per-warp timeline of a graph of chained compact step launches (CW_LIB_PATH -> the -DCW_TIMING build)."""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np, torch
import gym_craftingworld_b200 as cw
from gym_craftingworld_b200 import _lib
lib = _lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = 24
W = N // 32
env = cw.BatchedCraftingWorldEnv(N, seed=0, obs_mode="compact")
env.reset()
env.t.copy_(torch.randint(0, env.MAX_STEPS, (N,), device="cuda", dtype=torch.int32))
acts = torch.randint(0, 6, (K, N), device="cuda", dtype=torch.uint8)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    env.step(acts[0], chain_pos=0)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for k in range(K):
            env.step(acts[k], chain_pos=k)
    for _ in range(3):
        g.replay()
    s.synchronize()
    dbg = torch.zeros((K * W, 16), dtype=torch.int64, device="cuda")
    lib.cw_debug_set_timing.argtypes = [C.c_void_p]
    lib.cw_debug_set_timing_rows_per_position.argtypes = [C.c_int]
    assert lib.cw_debug_set_timing(dbg.data_ptr()) == 0 and lib.cw_debug_set_timing_rows_per_position(W) == 0
    g.replay()
    s.synchronize()
d = dbg.cpu().numpy().astype(np.float64).reshape(K, W, 16)
t0 = d[4, :, 0].min()
u = lambda x: (x - t0) / 1e3
print(f"chained compact step, {N} worlds = {W} warps/CTAs per launch; us relative to position 4's first entry")
print("pos | entry first / median / last | acquired median | stepped median | published median / last | exit last")
for p in range(4, K):
    e = d[p]
    print("%3d | %7.2f / %7.2f / %7.2f | %7.2f | %7.2f | %7.2f / %7.2f | %7.2f" % (
        p, u(e[:, 0].min()), u(np.median(e[:, 0])), u(e[:, 0].max()), u(np.median(e[:, 1])), u(np.median(e[:, 2])),
        u(np.median(e[:, 3])), u(e[:, 3].max()), u(e[:, 5].max())))
