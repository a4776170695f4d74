"""Interpreted restatement of the v0 step (test/bench infrastructure only).

The reference spends ~100 % of a step in a pure-Python per-pixel upsample loop
(reference gym_lmaze/envs/lmaze_env.py:219-234: 28,224 interpreted element writes,
~25 ms).  The C oracle removes that interpreter cost, so bench.py also times this
loop-for-loop Python restatement on a handful of steps to show, on the same host,
what the reference's own implementation style costs.  It is checked against the C
oracle in tests/test_oracle_golden.py::test_pyloop_matches_c_oracle.
"""
import numpy as np

REWARD = {"wall": -1.0, "move": -0.01, "goal": 100.0}


class PyLoopV0(object):
    def __init__(self, rows, spawn):
        self.rows = rows
        self.G, self.E = len(rows), 7
        G = self.G
        self.layers = np.zeros((4, G, G), dtype=np.float32)
        for i in range(G):
            for j in range(G):
                c = rows[i][j]
                self.layers[1, i, j] = 1.0 if c == "W" else 0.0     # lmaze_env.py:92-94
                self.layers[2, i, j] = 1.0 if c == "X" else 0.0     # :96-98
                self.layers[3, i, j] = 1.0 if c == "B" else 0.0     # :105-107
        self.goal_count = 0
        self.reset(spawn)

    def reset(self, spawn):
        self.layers[0] = 0.0
        self.x, self.y = spawn
        self.layers[0, self.x, self.y] = 1.0
        self.reward, self.step_count = -0.0, 0                       # :109-110
        return self._render()

    def _render(self):
        G, E = self.G, self.E
        out = np.zeros((4, G * E, G * E), dtype=np.float32)         # :217
        src = self.layers
        for c in range(4):                                           # :219-234, element by element
            for i in range(G):
                for ii in range(E):
                    for j in range(G):
                        for jj in range(E):
                            out[c][i * E + ii][j * E + jj] = src[c][i][j]
        return out

    def step(self, a):
        a = int(a)
        self.step_count += 1
        dx, dy = {0: (-1, 0), 1: (1, 0), 2: (0, -1), 3: (0, 1)}.get(a, (0, 0))   # :153-170
        t = self.rows[self.x + dx][self.y + dy]
        if t == "W":
            self.reward = REWARD["wall"]
        elif t in "BX":
            self.layers[0, self.x, self.y] = 0.0
            self.x += dx
            self.y += dy
            self.layers[0, self.x, self.y] = 1.0
            self.reward = REWARD["move"] if t == "B" else REWARD["goal"]
            self.goal_count += t == "X"
        done = self.reward == 100.0 or self.step_count == 100        # :246-249
        return self._render(), self.reward, done, a
