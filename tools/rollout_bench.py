#!/usr/bin/env python
"""Time the T-step rollout kernel (GPU box only):  python tools/rollout_bench.py [--envs N] [--T 64]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 21)
ap.add_argument("--T", type=int, default=64)
ap.add_argument("--variant", default="v0")
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
N, T = args.envs, args.T
env = lmz.LmazeVecCuda(N, args.variant, seed=3, with_obs=False)
env.reset()
rew = torch.empty((T, N), dtype=torch.float32, device="cuda")
don = torch.empty((T, N), dtype=torch.uint8, device="cuda")
acts = torch.randint(0, 4, (T, N), dtype=torch.uint8, device="cuda")
for name, a in (("device Philox actions", None), ("action buffer u8 [T,N]", acts)):
    for _ in range(3):
        env.rollout(T, actions=a, rewards=rew, dones=don)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        env.rollout(T, actions=a, rewards=rew, dones=don)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print("%s %s N=%d T=%d: %.3f ms/rollout  %.3e env-steps/s  (%.1f GB/s of reward+done)"
          % (args.variant, name, N, T, ms, N * T / ms * 1e3, N * T * 5 / ms / 1e6), flush=True)
print(env.stats())
