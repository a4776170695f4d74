"""Minimal stand-in for the `gym` package (which is not installed in this image).

Test infrastructure only.  The unmodified reference modules touch gym solely for
the `gym.Env` base class and the `spaces.Discrete` / `spaces.Box` declarations
(reference gym_lmaze/envs/lmaze_env.py:1-3,11,16,20), so this stub is enough to
execute them by file path.  Nothing in the product imports it.
"""
from . import spaces, error, utils  # noqa: F401


class Env(object):
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass
