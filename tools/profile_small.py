#!/usr/bin/env python
"""compact + incremental kernels at scale (for ncu; GPU box only)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
n = 1 << 22
env = lmz.LmazeVecCuda(n, "v0", seed=1, obs_mode="compact")
env.reset()
a = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8)
for _ in range(3):
    env.step(a)
torch.cuda.synchronize(); env.close(); del env
n = 1 << 19
env = lmz.LmazeVecCuda(n, "v0", seed=1, render_mode="incremental")
env.reset()
a = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8)
for _ in range(4):
    env.step(a)
torch.cuda.synchronize(); env.close()
print("ok")
