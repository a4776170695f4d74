import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        super().__init__((), np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = low
        self.high = high
        super().__init__(shape, dtype)
