#!/usr/bin/env python
"""One reset + a few steps of each foveal variant (v2, v4, v5) at a representative size (for ncu; GPU box only).
lmz_env_fov_kernel launches in order: v2 reset, 3 steps | v4 reset, 3 steps | v5 reset, (plannerStep, step) x 2."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
for variant, n in (("v2", 1 << 20), ("v4", 1 << 19)):
    env = lmz.LmazeVecCuda(n, variant, seed=1)
    env.reset()
    a = torch.randint(0, 25, (n,), device="cuda", dtype=torch.uint8)
    for _ in range(3):
        env.step(a)
    torch.cuda.synchronize(); env.close(); del env
n = 1 << 19
env = lmz.LmazeHierCuda(n, "v5", seed=1)
env.reset()
a = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8)
g = torch.randint(0, 25, (n,), device="cuda", dtype=torch.uint8)
for _ in range(2):
    env.plannerStep(g, mask="auto"); env.step(a, goal_plane=False)
torch.cuda.synchronize(); env.close()
print("ok")
