#!/usr/bin/env python
"""Golden fixtures for lmaze-v4 (v2 + float "visit map" layer), from the UNMODIFIED reference.

    python tests/golden/gen_golden_v4.py        (build container only; ~1 minute)

v4 differences that matter (reference gym_lmaze/envs/lmaze_env_v4.py):
  * reset(): the maze is re-rolled FIRST, goal/ball are then drawn on the NEW maze (:91-98)
  * state[2] = (state[2] + visitMap) / 2 every reset/step, in float64 then stored to float32
    (:116-119, :211-214); the 5x5 window around the ball is the visitMap
  * retStatelast is a VIEW of `state`, so the "previous crop" channels show the CURRENT
    state[2] values at the PREVIOUS window position (:122, :258-259)
  * obs (7,35,35): crop(3) + action one-hot + previous crop(3)

Outputs: v4_table.npz (state -> step -> outputs incl. the whole visit layer), v4_traces.npz.
The visit channels are stored un-expanded (5x5 float32); the generator asserts that the
reference obs is exactly their x7 replication.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.ref_loader import load_reference_module, ScriptedRandom  # noqa: E402
from gen_golden_v2 import ScriptedNumpy  # noqa: E402

BIT_CH = (0, 1, 3, 4, 5)      # free, goal, action, prev free, prev goal
VIS_CH = (2, 6)               # visit crop, previous-position visit crop


def split_obs(obs):
    """(7,35,35) f32 -> (packed bits of the 5 binary channels, (2,5,5) f32 visit crops)."""
    assert obs.dtype == np.float32 and obs.shape == (7, 35, 35)
    small = obs[:, ::7, ::7]
    assert np.array_equal(np.repeat(np.repeat(small, 7, 1), 7, 2), obs)
    b = small[list(BIT_CH)]
    assert ((b == 0) | (b == 1)).all()
    return np.packbits(b.astype(np.uint8).reshape(-1)), small[list(VIS_CH)].copy()


def make_env(layout, goal, ball):
    mod = load_reference_module("v4")
    sr, snp = ScriptedRandom(), ScriptedNumpy()
    mod.random, mod.np = sr, snp
    snp.random.queue.append(layout)
    sr.push(*goal); sr.push(*ball)
    with contextlib.redirect_stdout(io.StringIO()):
        env = mod.LmazeEnv_v4()
    return env, sr, snp


def main():
    layouts = [[str(r) for r in rows] for rows in np.load(os.path.join(HERE, "v2_layouts.npz"))["layouts"]]
    rng = np.random.RandomState(44)
    recs, bits_l, vis_l, pre_visit, post_visit = [], [], [], [], []
    for L in range(1, 6):
        rows = layouts[L - 1]
        gcells = [(x, y) for x in range(1, 17) for y in range(1, 17) if rows[x][y] not in "WS"]
        bcells = [(x, y) for x in range(1, 17) for y in range(1, 17) if rows[x][y] not in "WX"]
        for rep in range(6):
            g = gcells[rng.randint(len(gcells))]
            b = g
            while b == g:
                b = bcells[rng.randint(len(bcells))]
            env, sr, snp = make_env(L, g, b)
            with contextlib.redirect_stdout(io.StringIO()):
                for _ in range(rng.randint(0, 12)):          # a short walk to get a non-trivial visit layer
                    obs, r, d, _ = env.step(int(rng.randint(25)))
                    if d:
                        break
            if env.stepCount > 40 or d:
                continue
            base = dict(ball=(env.ball_x0, env.ball_y0), visit=env.state[2].copy(), step=env.stepCount)
            for a in range(25):
                env.ball_x0, env.ball_y0 = base["ball"]
                env.state[2] = base["visit"]
                env.stepCount = base["step"]
                env.retStatelast = np.asarray(env.state[:, base["ball"][0] - 2:base["ball"][0] + 3,
                                                        base["ball"][1] - 2:base["ball"][1] + 3])
                with contextlib.redirect_stdout(io.StringIO()):
                    obs, r, d, info = env.step(a)
                bb, vv = split_obs(obs)
                recs.append((L, base["ball"][0], base["ball"][1], g[0], g[1], a, base["step"], env.ball_x0, env.ball_y0,
                             np.float64(r).view(np.int64), int(d), env.stepCount))
                bits_l.append(bb); vis_l.append(vv)
                pre_visit.append(base["visit"]); post_visit.append(env.state[2].copy())
    cols = "layout bx by gx gy action step_before nx ny reward_bits done step_after"
    np.savez_compressed(os.path.join(HERE, "v4_table.npz"), columns=np.array(cols.split()),
                        table=np.array(recs, np.int64), obs_bits=np.stack(bits_l), obs_visit=np.stack(vis_l),
                        pre_visit=np.stack(pre_visit), post_visit=np.stack(post_visit))

    # ---- traces with resets (maze re-rolled first, goal/ball drawn on the new maze)
    T, NE = 220, 4
    out = {"n_envs": np.int64(NE)}

    def draw():
        L = int(rng.randint(1, 6))
        rows = layouts[L - 1]
        gc = [(x, y) for x in range(1, 17) for y in range(1, 17) if rows[x][y] not in "WS"]
        bc = [(x, y) for x in range(1, 17) for y in range(1, 17) if rows[x][y] not in "WX"]
        g = gc[rng.randint(len(gc))]
        b = g
        while b == g:
            b = bc[rng.randint(len(bc))]
        return L, g, b
    for e in range(NE):
        L, g, b = draw()
        env, sr, snp = make_env(L, g, b)
        fb, fv = split_obs(env.retStateExpanded.copy())
        acts = rng.randint(0, 25, size=T)
        pos = np.zeros((T, 2), np.int64); rew = np.zeros(T, np.int64); done = np.zeros(T, np.uint8)
        spawn = -np.ones((T, 5), np.int64); bits_t, vis_t = [], []; visit_sum = np.zeros(T, np.float64)
        for t in range(T):
            dx, dy = env.goal_x - env.ball_x0, env.goal_y - env.ball_y0
            if rng.rand() < 0.15 and abs(dx) <= 2 and abs(dy) <= 2:
                acts[t] = (dx + 2) * 5 + (dy + 2)
            with contextlib.redirect_stdout(io.StringIO()):
                obs, r, d, _ = env.step(int(acts[t]))
            pos[t] = (env.ball_x0, env.ball_y0); rew[t] = np.float64(r).view(np.int64); done[t] = d
            if d:
                L2, g2, b2 = draw()
                snp.random.queue.append(L2); sr.push(*g2); sr.push(*b2)
                obs = env.reset()
                spawn[t] = b2 + g2 + (L2,)
            bb, vv = split_obs(obs)
            bits_t.append(bb); vis_t.append(vv)
            visit_sum[t] = float(env.state[2].astype(np.float64).sum())
        out["e%d_init" % e] = np.array(b + g + (L,), np.int64)
        out["e%d_first_bits" % e] = fb; out["e%d_first_visit" % e] = fv
        out["e%d_actions" % e] = acts.astype(np.int64); out["e%d_pos" % e] = pos
        out["e%d_reward_bits" % e] = rew; out["e%d_done" % e] = done; out["e%d_spawn" % e] = spawn
        out["e%d_obs_bits" % e] = np.stack(bits_t); out["e%d_obs_visit" % e] = np.stack(vis_t)
        out["e%d_visit_sum" % e] = visit_sum
        out["e%d_final_visit" % e] = env.state[2].copy()
    np.savez_compressed(os.path.join(HERE, "v4_traces.npz"), **out)
    print("v4 fixtures written: %d table rows" % len(recs))


if __name__ == "__main__":
    main()
