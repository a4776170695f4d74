#!/usr/bin/env python
"""Golden fixtures for lmaze-v2 (multi-layout foveal env), from the UNMODIFIED reference.

    python tests/golden/gen_golden_v2.py        (build container only; ~1 minute)

The reference re-rolls its maze with np.random.randint(1, 6) (lmaze_env_v2.py:306) and draws
goal / ball with random.randint (:281-282, :294-295).  Both sources are replaced by scripted
stand-ins on the loaded module object (the file itself is untouched), so every draw is known.

Outputs:
  v2_layouts.npz  the five 18x18 mazes
  v2_table.npz    transition table: layouts x ball cells x goals x all 25 actions (+ step-limit
                  and border-clamp cases) -> next ball, reward bits, done, rendered obs (bit-packed)
  v2_traces.npz   4 scripted traces of 220 steps with resets (goal/ball sampled on the OLD maze,
                  then the maze re-rolled, :90-92)
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.ref_loader import load_reference_module, ScriptedRandom  # noqa: E402


class ScriptedNumpy(object):
    """numpy with np.random.randint scripted (everything else delegates to numpy)."""

    class _R(object):
        def __init__(self):
            self.queue = []

        def randint(self, a, b=None):
            v = self.queue.pop(0)
            assert a <= v < b
            return v

        def random(self):
            return 1.0

    def __init__(self):
        self.random = ScriptedNumpy._R()

    def __getattr__(self, k):
        return getattr(np, k)


def pack_obs(obs):
    o8 = obs.astype(np.uint8)
    assert obs.dtype == np.float32 and (o8.astype(np.float32) == obs).all()
    return np.packbits(o8.reshape(-1))


def make_env(first_layout, goal, ball, next_layout):
    mod = load_reference_module("v2")
    sr, snp = ScriptedRandom(), ScriptedNumpy()
    mod.random, mod.np = sr, snp
    snp.random.queue += [first_layout, next_layout]       # __init__: setGrid(); reset(): ... setGrid()
    sr.push(*goal); sr.push(*ball)
    with contextlib.redirect_stdout(io.StringIO()):
        env = mod.LmazeEnv_v2()
    return env, sr, snp


def all_layouts():
    out = []
    for L in range(1, 6):
        env, sr, snp = make_env(L, (8, 8) if L != 5 else (8, 7), (4, 5), L)
        out.append(["".join(r) for r in env.grid])
    return out


def place(env, rows, ball, goal, prev, step_count):
    """Force (maze, ball, goal, previous-crop position, stepCount) by writing the attributes
    reset()/step() maintain (lmaze_env_v2.py:82-123,212-218)."""
    env.grid = np.array([list(r) for r in rows])
    env.state = np.zeros((2, 18, 18), dtype=np.float32)
    env.state[0] = [[1.0 if c in "BSX" else 0.0 for c in row] for row in rows]
    env.goal_x, env.goal_y = goal
    env.state[1][goal[0]][goal[1]] = 1.0
    env.ball_x0, env.ball_y0 = ball
    env.retStatelast = np.asarray(env.state[:, prev[0] - 2:prev[0] + 3, prev[1] - 2:prev[1] + 3])
    env.stepCount = step_count


def main():
    layouts = all_layouts()
    np.savez_compressed(os.path.join(HERE, "v2_layouts.npz"), layouts=np.array(layouts))
    env, sr, snp = make_env(1, (8, 8), (4, 5), 1)
    rng = np.random.RandomState(42)
    recs, obs_bits = [], []
    for L in range(5):
        rows = layouts[L]
        free = [(x, y) for x in range(18) for y in range(18) if rows[x][y] != "W"]
        balls = [free[i] for i in rng.choice(len(free), 7, replace=False)] + [(2, 2), (15, 15), (2, 9), (9, 15), (3, 14)]
        goals = [free[i] for i in rng.choice(len(free), 2, replace=False)] + [(5, 5)]   # (5,5) is a wall in some mazes
        for ball in balls:
            for goal in goals:
                # retStatelast is normally the crop at the ball's own cell (that is what step()/reset()
                # leave behind); every other goal uses an arbitrary cell to show the oracle does not care
                prev = ball if (len(recs) // 25) % 2 == 0 else free[rng.randint(len(free))]
                for a in range(25):
                    place(env, rows, ball, goal, prev, 7)
                    obs, r, d, info = env.step(a)
                    assert info == a
                    recs.append((L + 1, ball[0], ball[1], goal[0], goal[1], prev[0], prev[1], a, 7,
                                 env.ball_x0, env.ball_y0, np.float64(r).view(np.int64), int(d), env.stepCount))
                    obs_bits.append(pack_obs(obs))
    edge = []
    for sc in (49, 50, 51, 60):                                    # done when stepCount > 50 (:222)
        for a in (12, 0, 24):
            place(env, layouts[0], (6, 6), (13, 13), (6, 6), sc)
            obs, r, d, _ = env.step(a)
            edge.append((1, 6, 6, 13, 13, 6, 6, a, sc, env.ball_x0, env.ball_y0,
                         np.float64(r).view(np.int64), int(d), env.stepCount))
            obs_bits.append(pack_obs(obs))
    cols = "layout bx by gx gy px py action step_before nx ny reward_bits done step_after"
    np.savez_compressed(os.path.join(HERE, "v2_table.npz"), columns=np.array(cols.split()),
                        table=np.array(recs + edge, np.int64), obs=np.stack(obs_bits))

    # ---- traces with resets
    T, NE = 220, 4
    out = {"n_envs": np.int64(NE)}
    for e in range(NE):
        L0 = int(rng.randint(1, 6))

        def draw(rows):
            goals = [(x, y) for x in range(1, 17) for y in range(1, 17) if rows[x][y] not in "WS"]
            g = goals[rng.randint(len(goals))]
            balls = [(x, y) for x in range(1, 17) for y in range(1, 17) if rows[x][y] not in "WX" and (x, y) != g]
            return g, balls[rng.randint(len(balls))]
        g, b = draw(layouts[L0 - 1])
        L = int(rng.randint(1, 6))
        env, sr, snp = make_env(L0, g, b, L)
        # bias the walk toward the goal now and then so that episodes also end by reward
        acts = rng.randint(0, 25, size=T)
        pos = np.zeros((T, 2), np.int64); rew = np.zeros(T, np.int64); done = np.zeros(T, np.uint8)
        spawn = -np.ones((T, 5), np.int64); obs_l = []; stepc = np.zeros(T, np.int64)
        first = pack_obs(env.retStateExpanded.copy())
        for t in range(T):
            dx, dy = env.goal_x - env.ball_x0, env.goal_y - env.ball_y0
            if rng.rand() < 0.15 and abs(dx) <= 2 and abs(dy) <= 2:
                acts[t] = (dx + 2) * 5 + (dy + 2)
            obs, r, d, _ = env.step(int(acts[t]))
            pos[t] = (env.ball_x0, env.ball_y0); rew[t] = np.float64(r).view(np.int64); done[t] = d
            stepc[t] = env.stepCount
            if d:
                g2, b2 = draw(["".join(r_) for r_ in env.grid])      # sampled on the maze BEFORE the re-roll
                L2 = int(rng.randint(1, 6))
                sr.push(*g2); sr.push(*b2); snp.random.queue.append(L2)
                obs = env.reset()
                spawn[t] = b2 + g2 + (L2,)
            obs_l.append(pack_obs(obs))
        out["e%d_init" % e] = np.array(b + g + (L, L0), np.int64)     # bx, by, gx, gy, layout, layout-before
        out["e%d_first_obs" % e] = first
        out["e%d_actions" % e] = acts.astype(np.int64)
        out["e%d_pos" % e] = pos; out["e%d_reward_bits" % e] = rew; out["e%d_done" % e] = done
        out["e%d_spawn" % e] = spawn; out["e%d_obs" % e] = np.stack(obs_l); out["e%d_step_count" % e] = stepc
    np.savez_compressed(os.path.join(HERE, "v2_traces.npz"), **out)
    print("v2 fixtures written: %d table rows" % (len(recs) + len(edge)))


if __name__ == "__main__":
    main()
