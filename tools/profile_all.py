#!/usr/bin/env python
"""Launch every kernel of the package a few times at a representative size (for ncu; GPU box only)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz

def run(tag, **kw):
    n = kw.pop("n")
    env = lmz.LmazeVecCuda(n, seed=1, **kw)
    env.reset()
    a = torch.randint(0, env.num_actions, (n,), device="cuda", dtype=torch.uint8)
    for _ in range(3):
        env.step(a)
    torch.cuda.synchronize()
    print(tag, "ok", flush=True)
    return env

run("v0_st128", n=1 << 18, variant="v0", render_mode="st128").close()
run("v3_tma", n=1 << 19, variant="v3").close()
run("v2_fov", n=1 << 20, variant="v2").close()
run("v4_fov", n=1 << 19, variant="v4").close()
run("v0_compact", n=1 << 22, variant="v0", obs_mode="compact").close()
e = run("v0_rollout", n=1 << 21, variant="v0", with_obs=False)
for _ in range(3):
    e.rollout(64)
torch.cuda.synchronize()
e.close()
