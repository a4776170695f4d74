#!/usr/bin/env python
"""Golden fixtures for lmaze-v5 / v6 (two-level planner/actor env), from the UNMODIFIED reference.

    python tests/golden/gen_golden_v5.py        (build container only; ~1 minute)

Protocol (reference gym_lmaze/envs/lmaze_env_v5.py): reset() -> foveal obs (7,35,35);
plannerStep(goal25) -> local obs (4,35,35); step(action4) -> 8-tuple (foveal obs, local obs,
globalReward, originalReward, globalDone, localDone, fovealGoal, action).

The reference crashes by construction when the actor leaves the planner-time fovea on the +x / +y side:
the same condition that sets localDone (`ballNew > fovea_x1 + 2`, :237-238) makes
buildLocalObservation index a 5x5 array with 5 (:365).  On the -x / -y side the index goes negative and
numpy WRAPS it silently.  The traces below are produced by an actor that never takes the crashing move
(it re-draws the action), so they contain the wrap-around cases but no IndexError.

Outputs: v5_traces.npz -- per env an event list (0 reset, 1 plannerStep, 2 step) with every returned
value; the visit channels of the foveal obs are stored un-expanded (5x5 f32), the binary channels packed.
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
from gen_golden_v4 import split_obs  # noqa: E402

DELTA = {0: (1, 0), 1: (-1, 0), 2: (0, 1), 3: (0, -1)}      # lmaze_env_v5.py:205-217


def pack_loc(obs):
    assert obs.dtype == np.float32 and obs.shape == (4, 35, 35)
    small = obs[:, ::7, ::7]
    assert np.array_equal(np.repeat(np.repeat(small, 7, 1), 7, 2), obs) and ((small == 0) | (small == 1)).all()
    return np.packbits(small.astype(np.uint8).reshape(-1))


def main(variant="v5"):
    layouts = [[str(r) for r in rows] for rows in np.load(os.path.join(HERE, "v2_layouts.npz"))["layouts"]]
    rng = np.random.RandomState(55)

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
    out = {"n_envs": np.int64(4)}
    T = 650
    for e in range(4):
        mod = load_reference_module(variant)
        sr, snp = ScriptedRandom(), ScriptedNumpy()
        mod.random, mod.np = sr, snp
        L, g, b = draw()
        snp.random.queue.append(L); sr.push(*g); sr.push(*b)
        with contextlib.redirect_stdout(io.StringIO()):
            env = getattr(mod, "LmazeEnv_" + variant)()
        ev = []          # kind, arg, gR bits, oR bits, gDone, lDone, x, y, spawn(5)
        fov_bits, fov_vis, loc_bits, visit_sum = [], [], [], []

        def rec(kind, arg, fov=None, loc=None, gr=-0.0, orr=-0.0, gd=0, ld=0, spawn=(-1,) * 5):
            ev.append((kind, arg, np.float64(gr).view(np.int64), np.float64(orr).view(np.int64), int(gd), int(ld),
                       env.ball_x0, env.ball_y0) + tuple(spawn))
            if fov is not None:
                bb, vv = split_obs(fov)
            else:
                bb, vv = np.zeros(16, np.uint8), np.zeros((2, 5, 5), np.float32)
            fov_bits.append(bb); fov_vis.append(vv)
            loc_bits.append(pack_loc(loc) if loc is not None else np.zeros(13, np.uint8))
            visit_sum.append(float(env.state[2].astype(np.float64).sum()))
        # the constructor's reset(): re-render through a fresh reset is not possible, so replay it
        snp.random.queue.append(L); sr.push(*g); sr.push(*b)
        rec(0, 0, fov=env.reset(), spawn=b + g + (L,))
        steps = 0
        while steps < T:
            # planner: mostly a free cell of the current fovea, sometimes anything (walls, far corners)
            goal25 = int(rng.randint(25))
            if rng.rand() < 0.7:
                for _ in range(10):
                    cx, cy = env.ball_x0 + goal25 // 5 - 2, env.ball_y0 + goal25 % 5 - 2
                    if env.grid[cx][cy] != "W":
                        break
                    goal25 = int(rng.randint(25))
            loc = env.plannerStep(goal25)
            rec(1, goal25, loc=loc)
            extra = int(rng.randint(0, 3)) if rng.rand() < 0.15 else 0     # sometimes keep stepping after localDone
            while True:
                for _try in range(20):
                    a = int(rng.randint(0, 4)) if rng.rand() > 0.06 else 7
                    dx, dy = DELTA.get(a, (0, 0))
                    nx, ny = env.ball_x0 + dx, env.ball_y0 + dy
                    moves = env.grid[nx][ny] != "W"
                    px, py = (nx, ny) if moves else (env.ball_x0, env.ball_y0)
                    if px - env.fovea_x1 + 2 <= 4 and py - env.fovea_y1 + 2 <= 4:
                        break                      # this move does not hit the reference's IndexError
                else:
                    a = 7
                with contextlib.redirect_stdout(io.StringIO()):
                    fov, loc, gr, orr, gd, ld, fg, act = env.step(a)
                assert act == a and fg.shape == (1, 5, 5)
                steps += 1
                spawn = (-1,) * 5
                if gd:
                    rec(2, a, fov=fov, loc=loc, gr=gr, orr=orr, gd=gd, ld=ld)
                    L2, g2, b2 = draw()
                    snp.random.queue.append(L2); sr.push(*g2); sr.push(*b2)
                    rec(0, 0, fov=env.reset(), spawn=b2 + g2 + (L2,))
                    break
                rec(2, a, fov=fov, loc=loc, gr=gr, orr=orr, gd=gd, ld=ld, spawn=spawn)
                if ld:
                    if extra == 0:
                        break
                    extra -= 1
        out["e%d_events" % e] = np.array(ev, np.int64)
        out["e%d_fov_bits" % e] = np.stack(fov_bits); out["e%d_fov_visit" % e] = np.stack(fov_vis)
        out["e%d_loc_bits" % e] = np.stack(loc_bits); out["e%d_visit_sum" % e] = np.array(visit_sum)
    # v6: safeFovealGoal() -- uniform re-draw until the 5x5 window cell is not a wall (lmaze_env_v6.py:505-523)
    if variant == "v6":
        draws = [int(v) for v in rng.randint(0, 25, size=400)]
        snp.random.queue.extend(draws)
        res = []
        n0 = len(snp.random.queue)
        for _ in range(40):
            before = len(snp.random.queue)
            res.append((env.ball_x0, env.ball_y0, env.safeFovealGoal(), before - len(snp.random.queue)))
        out["safe_layout"] = np.array(["".join(r) for r in env.grid])
        out["safe_draws"] = np.array(draws, np.int64); out["safe_results"] = np.array(res, np.int64)
    np.savez_compressed(os.path.join(HERE, variant + "_traces.npz"), **out)
    ev_all = np.concatenate([out["e%d_events" % e] for e in range(4)])
    print(variant, "events:", len(ev_all), "resets", int((ev_all[:, 0] == 0).sum()), "planner", int((ev_all[:, 0] == 1).sum()),
          "steps", int((ev_all[:, 0] == 2).sum()), "global dones", int(ev_all[:, 4].sum()))


if __name__ == "__main__":
    main("v5")
    main("v6")
