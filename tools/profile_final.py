#!/usr/bin/env python
"""One launch sequence over every kernel of the package at a representative size (for ncu; GPU box only).
Order of `lmz_*` launches (see profiles/README.md): each env is reset once, then stepped a few times."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz

def run(tag, n, variant, steps=1, **kw):
    env = lmz.LmazeVecCuda(n, variant, seed=1, **kw)
    env.reset()
    a = torch.randint(0, env.num_actions, (n,), device="cuda", dtype=torch.uint8)
    for _ in range(steps):
        env.step(a)
    torch.cuda.synchronize()
    print(tag, flush=True)
    return env

run("v0 tma", 1 << 19, "v0").close()
run("v0 st128", 1 << 18, "v0", render_mode="st128").close()
run("v3 tma", 1 << 19, "v3").close()
run("v2 fov", 1 << 20, "v2").close()
run("v4 fov", 1 << 19, "v4").close()
n = 1 << 19
h = lmz.LmazeHierCuda(n, "v5", seed=1)
h.reset()
a = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8)
g = torch.randint(0, 25, (n,), device="cuda", dtype=torch.uint8)
for _ in range(2):      # the second plannerStep has the realistic (small) mask
    h.plannerStep(g, mask="auto"); h.step(a, goal_plane=False)
torch.cuda.synchronize(); h.close(); print("v5", flush=True)
run("v0 compact", 1 << 22, "v0", obs_mode="compact").close()
run("v0 incremental", 1 << 19, "v0", steps=2, render_mode="incremental").close()
run("v2 compact", 1 << 22, "v2", obs_mode="compact").close()
run("v4 compact", 1 << 20, "v4", obs_mode="compact").close()
e = run("v0 transition-only", 1 << 22, "v0", with_obs=False)
e.rollout(64)
torch.cuda.synchronize(); e.close(); print("rollout", flush=True)
