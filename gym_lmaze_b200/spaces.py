"""Minimal gym-style spaces, used when neither `gym` nor `gymnasium` is installed.

Same attributes the reference declares (lmaze_env.py:16,20): `Discrete(n)` with
`.n`, `Box(low, high, shape)` with `.low/.high/.shape/.dtype`.
"""
import numpy as np


class Discrete(object):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def contains(self, x):
        return 0 <= int(x) < self.n

    def sample(self):
        return int(np.random.randint(self.n))

    def __repr__(self):
        return "Discrete(%d)" % self.n

    def __eq__(self, other):
        return getattr(other, "n", None) == self.n


class Box(object):
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high = float(low), float(high)
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(((x >= self.low) & (x <= self.high)).all())

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low, self.high, self.shape, self.dtype)

    def __eq__(self, other):
        return (getattr(other, "shape", None) == self.shape and getattr(other, "low", None) == self.low
                and getattr(other, "high", None) == self.high)


def make_spaces(n_actions, obs_shape):
    """Prefer real gym / gymnasium space classes when importable."""
    for modname in ("gymnasium", "gym"):
        try:
            mod = __import__(modname)
            if not hasattr(mod, "__version__"):   # a test stub, not the real package
                continue
            return (mod.spaces.Discrete(n_actions),
                    mod.spaces.Box(0.0, 1.0, shape=obs_shape, dtype=np.float32))
        except Exception:
            continue
    return Discrete(n_actions), Box(0.0, 1.0, obs_shape, np.float32)
