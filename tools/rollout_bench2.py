#!/usr/bin/env python
"""T=64 rollouts of every variant (GPU box only): ms per launch, env-steps/s."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
T = 64
for variant, n in (("v0", 1 << 21), ("v3", 1 << 21), ("v2", 1 << 21), ("v4", 1 << 21), ("v5", 1 << 20)):
    hier = variant == "v5"
    env = lmz.LmazeHierCuda(n, "v5", seed=1, obs_mode="compact") if hier else lmz.LmazeVecCuda(n, variant, seed=1, with_obs=False)
    env.reset()
    for _ in range(2):
        env.rollout(T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        env.rollout(T)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%s rollout T=%d n=%d: %.3f ms  %.2f G env-steps/s" % (variant, T, n, ms, n * T / ms / 1e6), flush=True)
    env.close()
