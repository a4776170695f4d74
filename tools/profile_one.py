#!/usr/bin/env python
"""A few steps of ONE configuration, for ncu (GPU box only).
    python tools/profile_one.py VARIANT [obs_mode] [log2 envs] [render_mode] [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
variant = sys.argv[1]
obs_mode = sys.argv[2] if len(sys.argv) > 2 else "full"
n = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 19)
render_mode = sys.argv[4] if len(sys.argv) > 4 else "tma"
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3      # ncu: -s (launches to skip) chooses which of them is captured
hier = variant == "v5"
env = (lmz.LmazeHierCuda(n, "v5", seed=1, obs_mode=obs_mode) if hier
       else lmz.LmazeVecCuda(n, variant, seed=1, obs_mode=obs_mode, render_mode=render_mode))
env.reset()
a = torch.randint(0, 4 if hier else env.num_actions, (4, n), device="cuda", dtype=torch.uint8)
g = torch.randint(0, 25, (4, n), device="cuda", dtype=torch.uint8)
for i in range(steps):
    if hier:
        env.plannerStep(g[i % 4], mask="auto"); env.step(a[i % 4], goal_plane=False)
    else:
        env.step(a[i % 4])
torch.cuda.synchronize()
print("done", variant, obs_mode, n, flush=True)
