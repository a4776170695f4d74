#!/usr/bin/env python
"""The reference's usage loop, on the B200 path.

    python examples/random_agent.py            (needs a B200; build first: python -m gym_lmaze_b200.build)

(1) drop-in single maze -- same calls and return types as gkm2708/gym-lmaze's LmazeEnv;
(2) the same loop over 65,536 mazes with the observation tensor staying on the GPU.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmaze  # noqa: E402

# (1) reference-style: obs is a numpy (4,84,84) float32 array, reward a float, done a bool, info the action
env = lmaze.make("lmaze-v0")
obs = env.reset()
ret, steps = 0.0, 0
for _ in range(300):
    a = env.action_space.sample()
    obs, r, d, info = env.step(a)
    ret += r
    steps += 1
    if d:
        obs = env.reset()
print("single maze: %d steps, return %.2f, obs %s %s" % (steps, ret, obs.shape, obs.dtype))
env.close()

# (2) vectorised: tensors live on the device; finished envs restart inside the same step
N = 1 << 16
vec = lmaze.make("lmaze-vec-v0", num_envs=N, seed=0, render_mode="incremental")
obs = vec.reset()                                   # float32 [N,4,84,84] CUDA tensor, updated in place
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    actions = torch.randint(0, 4, (N,), device=obs.device, dtype=torch.uint8)   # a policy(obs) would go here
    obs, reward, done, info = vec.step(actions)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("vectorised: %d mazes x 200 steps in %.3f s = %.1f M env-steps/s; stats %s"
      % (N, dt, N * 200 / dt / 1e6, vec.stats()))
vec.close()
