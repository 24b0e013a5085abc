"""Open-loop compact rollouts: K steps per launch (cw_rollout) vs one launch per step, 65536 worlds."""
import sys, torch
import gym_craftingworld_b200 as cw
N, K = 65536, 128
env = cw.BatchedCraftingWorldEnv(N, seed=0, obs_mode="compact")
env.reset()
tape = torch.randint(0, 6, (K, N), device="cuda", dtype=torch.uint8)
for _ in range(3): env.rollout(tape, return_trace=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps): env.rollout(tape, return_trace=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("cw_rollout K=128: %.2f us per env-step launch-equivalent, %.2f G env-steps/s" % (ms * 1e3 / (reps * K), N * K * reps / ms / 1e6))
rew, dn = env.rollout(tape)
e0.record()
for _ in range(reps): env.rollout(tape)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("cw_rollout K=128 with reward/done trace: %.2f G env-steps/s" % (N * K * reps / ms / 1e6))
