"""Small drivers for ncu captures: each target launches ONE kernel family a few dozen times (eager launches, no graphs).
    python tools/ncu_targets.py chained_cfg2|chained_cfg5|delta|incremental|compact|compact_chained|closed|pipe  [steps]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import gym_craftingworld_b200 as cw

target = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = "cuda"


def stagger(env):
    env.t.copy_(torch.randint(0, env.MAX_STEPS, (env.num_envs,), device=dev, dtype=torch.int32))


if target in ("chained_cfg2", "chained_cfg5", "closed"):
    N, size, ring = (16384, 32, 2) if target == "chained_cfg5" else (4096, 21, 4)
    env = cw.BatchedCraftingWorldEnv(N, size=(size, size), seed=0, obs_buffers=ring, auto_reset=target != "chained_cfg5")
    env.reset()
    if target == "chained_cfg5":
        import bench
        bench.dense_worlds(env, torch, 99)
    else:
        stagger(env)
    tape = torch.randint(0, 6, (steps, N), device=dev, dtype=torch.uint8)
    abuf = torch.zeros(N, dtype=torch.uint8, device=dev)
    for k in range(steps):
        if target == "closed":
            env.frame_policy(out=abuf)
            env.step(abuf)
        else:
            env.step(tape[k], chain_pos=k)
elif target == "delta":
    env = cw.HostCraftingWorldEnv(4096, seed=0, transport="delta")
    env.reset()
    env.load_state(t=np.random.RandomState(1).randint(0, 300, 4096))
    acts = np.random.RandomState(0).randint(0, 6, (steps, 4096)).astype(np.uint8)
    for k in range(steps):
        env.step(acts[k])
    env.close()
elif target == "incremental":
    N = 65536
    env = cw.BatchedCraftingWorldEnv(N, seed=0, render="incremental")
    env.reset(); stagger(env)
    tape = torch.randint(0, 6, (steps, N), device=dev, dtype=torch.uint8)
    for k in range(steps):
        env.step(tape[k])
elif target == "pipe":                                # the two-launch pipeline of the host-driven single step
    env = cw.HostCraftingWorldEnv(4096, seed=0, return_frames=False)
    env.reset()
    env.load_state(t=np.random.RandomState(1).randint(0, 300, 4096))
    acts = np.random.RandomState(0).randint(0, 6, (steps, 4096)).astype(np.uint8)
    for k in range(steps):
        env.step(acts[k])
    env.sync()
    env.close()
elif target == "compact_chained":
    N = 65536
    env = cw.BatchedCraftingWorldEnv(N, seed=0, obs_mode="compact")
    env.reset(); stagger(env)
    tape = torch.randint(0, 6, (steps, N), device=dev, dtype=torch.uint8)
    for k in range(steps):
        env.step(tape[k], chain_pos=k)
elif target == "compact":
    N = 65536
    env = cw.BatchedCraftingWorldEnv(N, seed=0, obs_mode="compact")
    env.reset(); stagger(env)
    tape = torch.randint(0, 6, (steps, N), device=dev, dtype=torch.uint8)
    for k in range(steps):
        env.step(tape[k])
else:
    raise SystemExit("unknown target " + target)
torch.cuda.synchronize()
print("ok", target)
