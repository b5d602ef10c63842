"""Scheduler diagnostics of the step kernel: scheduler iterations and full-physics passes per slab (stats[6], [7])."""
import sys, torch
import blockpuzzle_gym_b200 as bpg
B, K = 1 << 20, 64
env = bpg.make_vec("BlocksTouch-v0", B, device=0, seed=0)
env.reset()
a = torch.rand(K, B, 4, device="cuda") * 2 - 1
out = {}
for it in range(3):
    env.stats_reset()
    env.step_fused(a, auto_reset=True, out=out)
    torch.cuda.synchronize()
v = env.stats_tensor().cpu().numpy()
slabs = B / 128
print("iters/slab %.1f passes/slab %.1f worker_steps %.4f of steps; fill %.1f lanes/pass" % (v[6] / slabs, v[7] / slabs, v[5] / v[2], v[5] / max(v[7], 1)))
