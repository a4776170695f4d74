#!/usr/bin/env python
"""The planner / actor loop of lmaze-v5 / lmaze-v6 (reference gym_lmaze/envs/lmaze_env_v5.py), on the B200 path.

    python examples/hier_agent.py              (needs a B200; build first: python -m gym_lmaze_b200.build)

(1) drop-in single maze with the reference's calls: reset() / plannerStep(goal) / step(action) -> 8-tuple;
(2) the same protocol over 262,144 mazes: the planner acts where the local episode is over (device-side mask),
    the actor steps every maze, finished mazes restart inside the same step.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmaze  # noqa: E402

# (1) reference-style (lmaze-v6 = v5 + safeFovealGoal)
env = lmaze.make("lmaze-v6")
fov = env.reset()                                   # numpy (7,35,35) float32
steps, crashes = 0, 0
for episode in range(3):
    fov, global_done = env.reset(), False
    while not global_done:
        loc = env.plannerStep(env.safeFovealGoal())  # numpy (4,35,35): the actor's view
        local_done = False
        while not (local_done or global_done):
            try:
                fov, loc, g_r, l_r, global_done, local_done, goal_plane, a = env.step(steps % 4)
            except IndexError:                      # the reference raises here too (lmaze_env_v5.py:365-366)
                crashes += 1
                local_done = True
            steps += 1
print("single maze: %d actor steps, %d IndexError rows (actor left the planner-time fovea on the +x/+y side)"
      % (steps, crashes))
env.close()

# (2) vectorised
N = 1 << 18
vec = lmaze.make("lmaze-vec-v6", num_envs=N, seed=0)
fov = vec.reset()                                   # float32 [N,7,35,35] CUDA tensor, updated in place
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    goals = vec.safeFovealGoal()                    # a planner(fov) would go here
    loc = vec.plannerStep(goals, mask="auto")       # only the mazes waiting for their planner take the goal
    actions = torch.randint(0, 4, (N,), device=fov.device, dtype=torch.uint8)    # an actor(loc) would go here
    fov, loc, g_reward, l_reward, g_done, l_done, goal_plane, _ = vec.step(actions, goal_plane=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("vectorised: %d mazes x 100 planner/actor steps in %.3f s = %.1f M env-steps/s; stats %s"
      % (N, dt, N * 100 / dt / 1e6, vec.stats(check_errors=False)))
vec.close()
