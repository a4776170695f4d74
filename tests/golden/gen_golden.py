#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden.py

The reference modules are executed by file path under tests/_gymstub (see
oracle/ref_loader.py); nothing is copied from them.  Takes ~3 minutes because
every reference step() costs 15-25 ms (lmaze_env.py:219-234).

Outputs (all np.savez_compressed):
  v0_table.npz   exhaustive v0 transition table: 72 positions x 5 actions x 4
                 previous rewards (+ step-count edge cases) -> next position,
                 reward (f64), done, goalCount delta; and the rendered obs for
                 each of the 72 positions (bit-packed, values are exactly 0/1).
  v0_traces.npz  scripted-spawn traces (4 envs x 260 steps, invalid actions
                 sprinkled in) with per-step position, reward, done, spawn and
                 bit-packed obs; plus one 600-step trace driven by the real
                 stdlib `random` (random.seed(0)) as in BASELINE config 1.
  v3_table.npz   v3 transition table for 6 goals x all ball cells x 5 action
                 spellings, step-limit edge cases, and renders.
  v3_traces.npz  scripted-spawn v3 traces (4 envs x 260 steps).
  layouts.npz    the mazes as extracted from the reference + md5 of the rows.
"""
import hashlib
import os
import random as std_random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.ref_loader import make_reference_env, load_reference_module  # noqa: E402

INVALID = 7


def pack_obs(obs):
    o8 = obs.astype(np.uint8)
    assert obs.dtype == np.float32 and (o8.astype(np.float32) == obs).all()
    return np.packbits(o8.reshape(-1))


def grid_rows(env):
    return ["".join(r) for r in env.grid]


# ------------------------------------------------------------------ v0
def v0_place(env, x, y, reward, step_count):
    """Force the reference instance into (pos, previous reward, stepCount) by
    writing the same attributes step() mutates (lmaze_env.py:176-195)."""
    G = env.realgrid
    env.state[0:G * G] = 0.0
    env.ball_x0, env.ball_y0 = x, y
    env.state[x * G + y] = 1.0
    env.reward = reward
    env.stepCount = step_count


def gen_v0_table():
    env, scripted, _ = make_reference_env("v0", first_draws=(1, 1))
    rows = grid_rows(env)
    G = 12
    positions = [(x, y) for x in range(G) for y in range(G) if rows[x][y] != "W"]
    assert len(positions) == 72
    prev_rewards = [-0.0, -1.0, -0.01, 100.0]
    actions = [0, 1, 2, 3, INVALID]
    recs = []
    renders = []
    for (x, y) in positions:
        for a in actions:
            for pr in prev_rewards:
                v0_place(env, x, y, pr, 10)
                gc0 = env.goalCount
                obs, r, d, info = env.step(a)
                assert info == a
                recs.append((x, y, a, np.float64(pr).view(np.int64), 10,
                             env.ball_x0, env.ball_y0, np.float64(r).view(np.int64), int(d),
                             env.goalCount - gc0, env.stepCount))
        # obs is a pure function of the position: render via an invalid action on a fresh placement
        v0_place(env, x, y, -0.0, 10)
        # an invalid action never moves the ball (offset 0,0): target is the own cell
        obs, _, _, _ = env.step(INVALID)
        assert (env.ball_x0, env.ball_y0) == (x, y)
        renders.append(pack_obs(obs))
    # step-count edge cases (Q4): done exactly at stepCount == 100, not at 101
    edge = []
    for sc in (98, 99, 100, 101, 150):
        for (x, y, a) in ((3, 3, 1), (1, 2, 3), (1, 1, INVALID), (4, 5, 1)):
            v0_place(env, x, y, -0.01, sc)
            gc0 = env.goalCount
            obs, r, d, _ = env.step(a)
            edge.append((x, y, a, np.float64(-0.01).view(np.int64), sc,
                         env.ball_x0, env.ball_y0, np.float64(r).view(np.int64), int(d),
                         env.goalCount - gc0, env.stepCount))
    # reset(): reward -0.0, stepCount 0, goalCount kept (Q2, Q5)
    env.goalCount = 5
    scripted.push(3, 1)
    obs = env.reset()
    reset_info = np.array([env.ball_x0, env.ball_y0, np.float64(env.reward).view(np.int64),
                           env.stepCount, env.goalCount, env.goal_x, env.goal_y], dtype=np.int64)
    reset_obs = pack_obs(obs)
    # rejection loop: a rejected pair (wall), a rejected goal cell, then an accepted pair
    scripted.push(2, 2, 5, 5, 10, 10)
    env.reset()
    reject_info = np.array([env.ball_x0, env.ball_y0, scripted.calls], dtype=np.int64)
    # RANDOM_BALL = False spawns on the S cell (lmaze_env.py:82-89)
    env.RANDOM_BALL = False
    obs = env.reset()
    fixed_info = np.array([env.ball_x0, env.ball_y0], dtype=np.int64)
    fixed_obs = pack_obs(obs)
    cols = "x y action prev_reward_bits step_before nx ny reward_bits done goal_delta step_after"
    np.savez_compressed(
        os.path.join(HERE, "v0_table.npz"),
        columns=np.array(cols.split()), table=np.array(recs, dtype=np.int64),
        edge=np.array(edge, dtype=np.int64), positions=np.array(positions, dtype=np.int64),
        renders=np.stack(renders), reset_info=reset_info, reset_obs=reset_obs,
        reject_info=reject_info, fixed_info=fixed_info, fixed_obs=fixed_obs)
    return rows


def gen_v0_traces(rows):
    G = 12
    spawnable = [(x, y) for x in range(G) for y in range(G) if rows[x][y] in "BS"]
    assert len(spawnable) == 71
    T, NE = 260, 4
    rng = np.random.RandomState(20261018)
    out = {}
    for e in range(NE):
        sp = spawnable[rng.randint(71)]
        env, scripted, _ = make_reference_env("v0", first_draws=sp)
        acts = rng.randint(0, 4, size=T)
        acts[rng.rand(T) < 0.05] = INVALID
        if e == 0:
            sp = (1, 2)
            scripted.push(*sp); env.reset()
            acts[:6] = [2, 2, 0, 2, INVALID, 3]      # walk into S: stale reward (Q1)
        if e == 1:
            sp = (3, 5)                                # (3,5)->(4,5)->(5,5): goal within 2 steps
            scripted.push(*sp); env.reset()
            acts[:2] = [1, 1]
        pos = np.zeros((T, 2), np.int64); rew = np.zeros(T, np.int64); done = np.zeros(T, np.uint8)
        spawn = -np.ones((T, 2), np.int64); obs_bits = []
        stepc = np.zeros(T, np.int64); goalc = np.zeros(T, np.int64)
        for t in range(T):
            obs, r, d, _ = env.step(int(acts[t]))
            pos[t] = (env.ball_x0, env.ball_y0); rew[t] = np.float64(r).view(np.int64); done[t] = d
            stepc[t] = env.stepCount; goalc[t] = env.goalCount
            if d:
                s = spawnable[rng.randint(71)]
                scripted.push(*s)
                obs = env.reset()
                spawn[t] = s
            obs_bits.append(pack_obs(obs))     # what a same-step auto-reset env returns
        out["e%d_spawn0" % e] = np.array(sp, np.int64)
        out["e%d_actions" % e] = acts.astype(np.int64)
        out["e%d_pos" % e] = pos; out["e%d_reward_bits" % e] = rew; out["e%d_done" % e] = done
        out["e%d_spawn" % e] = spawn; out["e%d_obs" % e] = np.stack(obs_bits)
        out["e%d_step_count" % e] = stepc; out["e%d_goal_count" % e] = goalc
    # BASELINE config 1 style: the real stdlib RNG drives the spawns
    mod = load_reference_module("v0")
    std_random.seed(0)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        env = mod.LmazeEnv()
    arng = std_random.Random(1)
    T = 600
    acts = np.zeros(T, np.int64); pos = np.zeros((T, 2), np.int64); rew = np.zeros(T, np.int64)
    done = np.zeros(T, np.uint8); spawn = -np.ones((T, 2), np.int64)
    sp0 = (env.ball_x0, env.ball_y0)
    for t in range(T):
        a = arng.randrange(4)
        obs, r, d, _ = env.step(a)
        acts[t] = a; pos[t] = (env.ball_x0, env.ball_y0); rew[t] = np.float64(r).view(np.int64); done[t] = d
        if d:
            env.reset()
            spawn[t] = (env.ball_x0, env.ball_y0)
    out.update(mt_spawn0=np.array(sp0, np.int64), mt_actions=acts, mt_pos=pos, mt_reward_bits=rew,
               mt_done=done, mt_spawn=spawn, n_envs=np.int64(NE))
    np.savez_compressed(os.path.join(HERE, "v0_traces.npz"), **out)


# ------------------------------------------------------------------ v3
V3_SPELL = {0: ("left", "0"), 1: ("right", "1"), 2: ("up", "2"), 3: ("down", "3")}


def v3_action(code, k):
    """Map an int action code to one of the reference's accepted strings (lmaze_env_v3.py:236-247)."""
    if code in V3_SPELL:
        return V3_SPELL[code][k % 2]
    return ("noop", 0, 3, "4")[k % 4]    # ints never match a string compare -> offset (0,0)


def v3_place(env, bx, by, gx, gy, step_count):
    env.state[1, :, :] = 0.0
    env.ball_x0, env.ball_y0 = bx, by
    env.state[1, bx, by] = 1.0
    env.goal_x, env.goal_y = gx, gy
    env.image_global_goal[:, :] = 0.0
    env.image_global_goal[gx][gy] = 1.0
    env.stepCount = step_count


def gen_v3():
    env, scripted, _ = make_reference_env("v3", first_draws=(8, 8, 7, 8))
    rows = grid_rows(env)
    G = 18
    free = [(x, y) for x in range(G) for y in range(G) if rows[x][y] != "W"]
    assert len(free) == 73
    goals = [(8, 8), (4, 4), (13, 13), (6, 9), (10, 12), (13, 4)]
    recs, renders, render_keys = [], [], []
    k = 0
    for (gx, gy) in goals:
        for (bx, by) in free:            # includes ball standing on the goal
            for a in (0, 1, 2, 3, INVALID):
                v3_place(env, bx, by, gx, gy, 10)
                obs, r, d, info = env.step(v3_action(a, k)); k += 1
                recs.append((bx, by, gx, gy, a, 10, env.ball_x0, env.ball_y0,
                             np.float64(r).view(np.int64), int(d), env.stepCount))
            v3_place(env, bx, by, gx, gy, 10)
            # 3 never equals a string -> offset 0,0 -> position kept; obs = f(ball, goal)
            obs, _, _, _ = env.step(3)
            assert (env.ball_x0, env.ball_y0) == (bx, by)
            renders.append(pack_obs(obs)); render_keys.append((bx, by, gx, gy))
    edge = []
    for sc in (98, 99, 100, 101, 150):
        for (bx, by, a) in ((6, 6, 1), (4, 5, 3), (4, 4, INVALID), (13, 13, 1)):
            v3_place(env, bx, by, 8, 8, sc)
            obs, r, d, _ = env.step(v3_action(a, k)); k += 1
            edge.append((bx, by, 8, 8, a, sc, env.ball_x0, env.ball_y0,
                         np.float64(r).view(np.int64), int(d), env.stepCount))
    # reset(): goal pair then ball pair, with rejected draws in between
    scripted.push(1, 1, 6, 9)            # goal: wall rejected, then (6,9)
    scripted.push(6, 9, 2, 2, 7, 7)      # ball: goal cell rejected, wall rejected, then (7,7)
    obs = env.reset()
    reset_info = np.array([env.ball_x0, env.ball_y0, env.goal_x, env.goal_y, env.stepCount], np.int64)
    reset_obs = pack_obs(obs)
    obs_t = env.reset(mode="test")       # lmaze_env_v3.py:145-146,154-155
    test_info = np.array([env.ball_x0, env.ball_y0, env.goal_x, env.goal_y], np.int64)
    cols = "bx by gx gy action step_before nx ny reward_bits done step_after"
    np.savez_compressed(
        os.path.join(HERE, "v3_table.npz"), columns=np.array(cols.split()),
        table=np.array(recs, np.int64), edge=np.array(edge, np.int64), free=np.array(free, np.int64),
        renders=np.stack(renders), render_keys=np.array(render_keys, np.int64),
        reset_info=reset_info, reset_obs=reset_obs, test_info=test_info, test_obs=pack_obs(obs_t))

    # traces
    T, NE = 260, 4
    rng = np.random.RandomState(20261019)
    out = {}
    for e in range(NE):
        def draw():
            g = free[rng.randint(73)]
            while True:
                b = free[rng.randint(73)]
                if b != g:
                    return g, b
        g, b = draw()
        env, scripted, _ = make_reference_env("v3", first_draws=g + b)
        acts = rng.randint(0, 4, size=T)
        acts[rng.rand(T) < 0.05] = INVALID
        pos = np.zeros((T, 4), np.int64); rew = np.zeros(T, np.int64); done = np.zeros(T, np.uint8)
        spawn = -np.ones((T, 4), np.int64); obs_bits = []; stepc = np.zeros(T, np.int64)
        for t in range(T):
            obs, r, d, _ = env.step(v3_action(int(acts[t]), t))
            pos[t] = (env.ball_x0, env.ball_y0, env.goal_x, env.goal_y)
            rew[t] = np.float64(r).view(np.int64); done[t] = d; stepc[t] = env.stepCount
            if d:
                g2, b2 = draw()
                scripted.push(*(g2 + b2))
                obs = env.reset()
                spawn[t] = b2 + g2
            obs_bits.append(pack_obs(obs))
        out["e%d_spawn0" % e] = np.array(b + g, np.int64)      # sx, sy, gx, gy
        out["e%d_actions" % e] = acts.astype(np.int64)
        out["e%d_pos" % e] = pos; out["e%d_reward_bits" % e] = rew; out["e%d_done" % e] = done
        out["e%d_spawn" % e] = spawn; out["e%d_obs" % e] = np.stack(obs_bits)
        out["e%d_step_count" % e] = stepc
    out["n_envs"] = np.int64(NE)
    np.savez_compressed(os.path.join(HERE, "v3_traces.npz"), **out)
    return rows


def main():
    t0 = time.time()
    rows0 = gen_v0_table(); print("v0 table  %.0fs" % (time.time() - t0))
    gen_v0_traces(rows0); print("v0 traces %.0fs" % (time.time() - t0))
    rows3 = gen_v3(); print("v3        %.0fs" % (time.time() - t0))
    np.savez_compressed(
        os.path.join(HERE, "layouts.npz"), v0=np.array(rows0), v3=np.array(rows3),
        v0_md5=np.array(hashlib.md5("/".join(rows0).encode()).hexdigest()),
        v3_md5=np.array(hashlib.md5("/".join(rows3).encode()).hexdigest()))


if __name__ == "__main__":
    main()
