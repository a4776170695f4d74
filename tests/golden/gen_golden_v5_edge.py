#!/usr/bin/env python
"""Edge-case fixtures for lmaze-v5 (planner / actor env), from the UNMODIFIED reference.

    python tests/golden/gen_golden_v5_edge.py        (build container only)

Same event format as gen_golden_v5.py (0 reset, 1 plannerStep, 2 step; every returned value recorded), but the
scripts walk the paths the random traces never take:
  * step() right after reset(), before any plannerStep (fovealStepCount == 0: retStatelast is re-pointed on
    every step, lmaze_env_v5.py:327-328; f_goal is the spawn cell, :144-145);
  * stepping on after globalDone without a reset (goal hit, and the fovealStepCount >= 50 timeout, :260-262);
  * the global goal placed next to the ball, so that goal hits, "goal == foveal goal" and "goal one step away"
    all occur (:245-251);
  * plannerStep with each of the 25 cells, walls and cells outside the maze included (:165-171).
The actor never takes the move on which the reference itself raises IndexError (see gen_golden_v5.py).
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
from oracle.ref_loader import load_reference_module, ScriptedRandom  # noqa: E402
from gen_golden_v2 import ScriptedNumpy  # noqa: E402
from gen_golden_v4 import split_obs  # noqa: E402
from gen_golden_v5 import pack_loc, DELTA  # noqa: E402


def main():
    layouts = [[str(r) for r in rows] for rows in np.load(os.path.join(HERE, "v2_layouts.npz"))["layouts"]]
    rng = np.random.RandomState(77)
    out = {}
    n_envs = 0

    def free(rows, x, y):
        return rows[x][y] != "W"

    def run(script_name, L, goal, ball, script):
        """script(env, rec_planner, rec_step) drives one env after its constructor + explicit reset."""
        nonlocal n_envs
        mod = load_reference_module("v5")
        sr, snp = ScriptedRandom(), ScriptedNumpy()
        mod.random, mod.np = sr, snp
        snp.random.queue.append(L); sr.push(*goal); sr.push(*ball)
        with contextlib.redirect_stdout(io.StringIO()):
            env = mod.LmazeEnv_v5()
        ev, fov_bits, fov_vis, loc_bits, visit_sum = [], [], [], [], []

        def rec(kind, arg, fov=None, loc=None, gr=-0.0, orr=-0.0, gd=0, ld=0, spawn=(-1,) * 5):
            ev.append((kind, arg, np.float64(gr).view(np.int64), np.float64(orr).view(np.int64), int(gd), int(ld),
                       env.ball_x0, env.ball_y0) + tuple(spawn))
            bb, vv = split_obs(fov) if fov is not None else (np.zeros(16, np.uint8), np.zeros((2, 5, 5), np.float32))
            fov_bits.append(bb); fov_vis.append(vv)
            loc_bits.append(pack_loc(loc) if loc is not None else np.zeros(13, np.uint8))
            visit_sum.append(float(env.state[2].astype(np.float64).sum()))

        def do_reset(L2, g2, b2):
            snp.random.queue.append(L2); sr.push(*g2); sr.push(*b2)
            rec(0, 0, fov=env.reset(), spawn=tuple(b2) + tuple(g2) + (L2,))

        def do_plan(goal25):
            rec(1, goal25, loc=env.plannerStep(goal25))

        def safe(a):
            dx, dy = DELTA.get(a, (0, 0))
            nx, ny = env.ball_x0 + dx, env.ball_y0 + dy
            px, py = (nx, ny) if env.grid[nx][ny] != "W" else (env.ball_x0, env.ball_y0)
            return px - env.fovea_x1 + 2 <= 4 and py - env.fovea_y1 + 2 <= 4

        def do_step(a):
            if not safe(a):
                a = 7                                   # the reference would raise IndexError: stand still instead
            with contextlib.redirect_stdout(io.StringIO()):
                fov, loc, gr, orr, gd, ld, fg, act = env.step(a)
            rec(2, a, fov=fov, loc=loc, gr=gr, orr=orr, gd=gd, ld=ld)
            return gd, ld

        do_reset(L, goal, ball)
        script(env, do_reset, do_plan, do_step)
        e = n_envs
        out["e%d_events" % e] = np.array(ev, np.int64)
        out["e%d_fov_bits" % e] = np.stack(fov_bits); out["e%d_fov_visit" % e] = np.stack(fov_vis)
        out["e%d_loc_bits" % e] = np.stack(loc_bits); out["e%d_visit_sum" % e] = np.array(visit_sum)
        out["e%d_name" % e] = np.array(script_name)
        n_envs += 1

    def cells(rows, pred):
        return [(x, y) for x in range(1, 17) for y in range(1, 17) if pred(rows[x][y])]

    # (a) steps before any plannerStep, on every layout
    for L in range(1, 6):
        rows = layouts[L - 1]
        bc = cells(rows, lambda c: c not in "WX")
        gc = cells(rows, lambda c: c not in "WS")
        b = bc[rng.randint(len(bc))]
        g = [c for c in gc if c != b][rng.randint(len(gc) - 1)]

        def script(env, do_reset, do_plan, do_step):
            for _ in range(14):                         # stepCount passes 10 -> localDone without a planner
                do_step(int(rng.randint(0, 4)))
            do_plan(int(rng.randint(25)))
            for _ in range(6):
                do_step(int(rng.randint(0, 4)))
        run("steps_before_planner_L%d" % L, L, g, b, script)

    # (b) + (c) goal next to the ball: walk onto it, then keep stepping past globalDone
    for L in range(1, 6):
        rows = layouts[L - 1]
        pairs = []
        for (x, y) in cells(rows, lambda c: c not in "WX"):
            for a, (dx, dy) in DELTA.items():
                gx, gy = x + dx, y + dy
                if rows[gx][gy] not in "WS":
                    pairs.append(((x, y), (gx, gy), a))
        for k in range(3):
            b, g, a = pairs[rng.randint(len(pairs))]

            def script(env, do_reset, do_plan, do_step, a=a, k=k):
                do_plan([12 + {0: 5, 1: -5, 2: 1, 3: -1}[a], int(rng.randint(25)), 12][k])   # goal == / != the foveal goal
                gd, ld = do_step(a)                     # onto the global goal (unless the IndexError guard vetoes it)
                for _ in range(5):                      # on past globalDone
                    do_step(int(rng.randint(0, 4)))
                do_plan(int(rng.randint(25)))
                for _ in range(4):
                    do_step(int(rng.randint(0, 4)))
            run("goal_adjacent_L%d_%d" % (L, k), L, g, b, script)

    # (d) every planner cell, then the 50-planner-step timeout and beyond
    rows = layouts[2]
    bc = cells(rows, lambda c: c not in "WX")
    b = bc[rng.randint(len(bc))]
    g = (13, 13) if b != (13, 13) else (4, 13)

    def script(env, do_reset, do_plan, do_step):
        for goal25 in list(range(25)) + [int(v) for v in rng.randint(0, 25, size=30)]:   # 55 planner steps: past the limit of 50
            do_plan(goal25)
            for _ in range(int(rng.randint(1, 3))):
                gd, ld = do_step(int(rng.randint(0, 4)))
    run("all_planner_cells_and_timeout", 3, g, b, script)

    out["n_envs"] = np.int64(n_envs)
    np.savez_compressed(os.path.join(HERE, "v5_edge_traces.npz"), **out)
    ev_all = np.concatenate([out["e%d_events" % e] for e in range(n_envs)])
    print("envs", n_envs, "events", len(ev_all), "steps", int((ev_all[:, 0] == 2).sum()), "global-done steps",
          int(ev_all[ev_all[:, 0] == 2][:, 4].sum()), "goal rewards",
          int((ev_all[:, 2] == np.float64(100.0).view(np.int64)).sum()))


if __name__ == "__main__":
    main()
