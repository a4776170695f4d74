"""CPU tier, build container only: the C oracle stepped LIVE beside the unmodified reference on fresh random
traces (seeds that the committed golden fixtures never used).  Skipped where /root/reference does not exist
(the GPU box) -- there the committed fixtures of tests/golden/ carry the pin.
"""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference checkout not present")


def f32bits(x):
    return np.float32(x).view(np.uint32)


def _cells(rows, pred):
    return [(x, y) for x in range(1, len(rows) - 1) for y in range(1, len(rows) - 1) if pred(rows[x][y])]


@pytest.mark.parametrize("seed", [101, 202])
def test_v0_live(oracle_mod, seed):
    rng = np.random.RandomState(seed)
    rows = oracle_mod.layout(oracle_mod.V0)
    spawnable = _cells(rows, lambda c: c not in "WX")
    b = spawnable[rng.randint(len(spawnable))]
    env, scripted, _ = ref_loader.make_reference_env("v0", first_draws=b)
    ora = oracle_mod.OracleVec(oracle_mod.V0, 1, autoreset=False)
    scripted.push(*b)
    assert np.array_equal(env.reset(), ora.reset(spawn=[[b[0], b[1], -1, -1]])[0])
    n_done = 0
    for t in range(160):
        a = int(rng.randint(0, 4)) if rng.rand() > 0.08 else int(rng.choice([-1, 4, 9]))
        obs, r, d, info = env.step(a)
        o_ref, r_o, d_o = ora.step([a])
        assert np.array_equal(obs, o_ref[0]) and f32bits(r) == r_o.view(np.uint32)[0] and bool(d) == bool(d_o[0]), t
        assert ora.export()[3].view(np.int64)[0] == np.float64(r).view(np.int64)      # the oracle keeps the Python float
        if d or rng.rand() < 0.02:
            b = spawnable[rng.randint(len(spawnable))]
            scripted.push(*b)
            assert np.array_equal(env.reset(), ora.reset(spawn=[[b[0], b[1], -1, -1]])[0])
            n_done += 1
    assert n_done >= 1


@pytest.mark.parametrize("seed", [303])
def test_v3_live(oracle_mod, seed):
    rng = np.random.RandomState(seed)
    rows = oracle_mod.layout(oracle_mod.V3)
    free = _cells(rows, lambda c: c != "W")

    def draw():
        g = free[rng.randint(len(free))]
        b = g
        while b == g:
            b = free[rng.randint(len(free))]
        return g, b
    g, b = draw()
    env, scripted, _ = ref_loader.make_reference_env("v3", first_draws=g + b)
    ora = oracle_mod.OracleVec(oracle_mod.V3, 1, autoreset=False)
    scripted.push(*g); scripted.push(*b)
    assert np.array_equal(env.reset(), ora.reset(spawn=[[b[0], b[1], g[0], g[1]]])[0])
    words = ["left", "right", "up", "down", "0", "1", "2", "3", "stay", 2]
    code = {"left": 0, "0": 0, "right": 1, "1": 1, "up": 2, "2": 2, "down": 3, "3": 3}
    for t in range(260):
        w = words[rng.randint(len(words))]
        obs, r, d, info = env.step(w)
        o_ref, r_o, d_o = ora.step([code.get(w, 255) if isinstance(w, str) else 255])
        assert np.array_equal(obs, o_ref[0]) and f32bits(r) == r_o.view(np.uint32)[0] and bool(d) == bool(d_o[0]), t
        if d:
            g, b = draw()
            scripted.push(*g); scripted.push(*b)
            assert np.array_equal(env.reset(), ora.reset(spawn=[[b[0], b[1], g[0], g[1]]])[0])


@pytest.mark.parametrize("variant", ["v2", "v4"])
def test_foveal_live(oracle_mod, variant):
    from gen_golden_v2 import ScriptedNumpy
    rng = np.random.RandomState(404 if variant == "v2" else 505)
    layouts = [oracle_mod.layout_v2(k) for k in range(1, 6)]

    def draw(L):
        rows = layouts[L - 1]
        gc = _cells(rows, lambda c: c not in "WS")
        bc = _cells(rows, lambda c: c not in "WX")
        g = gc[rng.randint(len(gc))]
        b = g
        while b == g:
            b = bc[rng.randint(len(bc))]
        return g, b
    mod = ref_loader.load_reference_module(variant)
    sr, snp = ref_loader.ScriptedRandom(), ScriptedNumpy()
    mod.random, mod.np = sr, snp
    ov = oracle_mod.V2 if variant == "v2" else oracle_mod.V4
    ora = oracle_mod.OracleVec(ov, 1, autoreset=False)
    cur = int(rng.randint(1, 6))

    def queue_reset(cur):
        """v2: goal / ball are drawn on the CURRENT maze, then the maze is re-rolled; v4: re-roll first."""
        nl = int(rng.randint(1, 6))
        g, b = draw(cur if variant == "v2" else nl)
        snp.random.queue.append(nl); sr.push(*g); sr.push(*b)
        return g, b, nl
    # constructor: v2 rolls a maze and then runs a full reset(); v4 only runs the reset (its setGrid() call in
    # __init__ is commented out, lmaze_env_v4.py:81); then an explicit reset we can mirror
    if variant == "v2":
        snp.random.queue.append(cur)
    g, b, nl = queue_reset(cur)
    with contextlib.redirect_stdout(io.StringIO()):
        env = getattr(mod, "LmazeEnv_" + variant)()
    cur = nl
    ora.L.lmzo_env_force_v2(ora._env(0), cur, b[0], b[1], g[0], g[1], b[0], b[1], 0)     # put the oracle on the same maze
    g, b, nl = queue_reset(cur)
    obs = env.reset()
    o_ref = ora.reset(spawn=[[b[0], b[1], g[0], g[1] | (nl << 5)]])
    cur = nl
    assert np.array_equal(obs.view(np.uint32), o_ref[0].view(np.uint32))
    for t in range(220):
        a = int(rng.randint(0, 25))
        with contextlib.redirect_stdout(io.StringIO()):
            obs, r, d, info = env.step(a)
        o_ref, r_o, d_o = ora.step([a])
        assert np.array_equal(obs.view(np.uint32), o_ref[0].view(np.uint32)), t
        assert f32bits(r) == r_o.view(np.uint32)[0] and bool(d) == bool(d_o[0]), t
        if d:
            g, b, nl = queue_reset(cur)
            obs = env.reset()
            o_ref = ora.reset(spawn=[[b[0], b[1], g[0], g[1] | (nl << 5)]])
            cur = nl
            assert np.array_equal(obs.view(np.uint32), o_ref[0].view(np.uint32)), t
    if variant == "v4":
        assert np.array_equal(np.asarray(env.state[2], np.float32).view(np.uint32), ora.export_visit()[0].view(np.uint32))


@pytest.mark.parametrize("variant", ["v5", "v6"])
def test_hier_live(oracle_mod, variant):
    """reset / plannerStep / step of the planner-actor env, live, incl. steps before the first plannerStep,
    stepping past localDone and safeFovealGoal (v6)."""
    from gen_golden_v2 import ScriptedNumpy
    from gen_golden_v5 import DELTA
    rng = np.random.RandomState(606 if variant == "v5" else 707)
    layouts = [oracle_mod.layout_v2(k) for k in range(1, 6)]

    def draw():
        L = int(rng.randint(1, 6))
        rows = layouts[L - 1]
        gc = _cells(rows, lambda c: c not in "WS")
        bc = _cells(rows, lambda c: c not in "WX")
        g = gc[rng.randint(len(gc))]
        b = g
        while b == g:
            b = bc[rng.randint(len(bc))]
        return L, g, b
    mod = ref_loader.load_reference_module(variant)
    sr, snp = ref_loader.ScriptedRandom(), ScriptedNumpy()
    mod.random, mod.np = sr, snp
    L, g, b = draw()
    snp.random.queue.append(L); sr.push(*g); sr.push(*b)
    with contextlib.redirect_stdout(io.StringIO()):
        env = getattr(mod, "LmazeEnv_" + variant)()
    ora = oracle_mod.OracleHier(1)

    def reset_both():
        L, g, b = draw()
        snp.random.queue.append(L); sr.push(*g); sr.push(*b)
        fov = env.reset()
        f_o = ora.reset(spawn=[[b[0], b[1], g[0], g[1] | (L << 5)]])
        assert np.array_equal(fov.view(np.uint32), f_o[0].view(np.uint32))
    reset_both()
    steps = 0
    planned = False
    while steps < 260:
        if planned or rng.rand() < 0.7:                 # sometimes step before the first plannerStep of an episode
            if variant == "v6" and rng.rand() < 0.5:
                draws = [int(v) for v in rng.randint(0, 25, size=60)]
                snp.random.queue.extend(draws)
                goal25 = env.safeFovealGoal()
                used = 60 - len(snp.random.queue)
                assert ora.safe_goal(0, draws) == (goal25, used)
                del snp.random.queue[:]
            else:
                goal25 = int(rng.randint(25))
            loc = env.plannerStep(goal25)
            l_o, err = ora.planner_step([goal25])
            assert not err[0] and np.array_equal(loc, l_o[0]), steps
            planned = True
        for _ in range(int(rng.randint(1, 14))):
            for _try in range(20):
                a = int(rng.randint(0, 4)) if rng.rand() > 0.06 else 7
                dx, dy = DELTA.get(a, (0, 0))
                nx, ny = env.ball_x0 + dx, env.ball_y0 + dy
                px, py = (nx, ny) if env.grid[nx][ny] != "W" else (env.ball_x0, env.ball_y0)
                if px - env.fovea_x1 + 2 <= 4 and py - env.fovea_y1 + 2 <= 4:
                    break                               # a move on which the reference itself does not crash
            else:
                a = 7
            with contextlib.redirect_stdout(io.StringIO()):
                fov, loc, gr, orr, gd, ld, fg, act = env.step(a)
            f_o, l_o, gr_o, lr_o, gd_o, ld_o, err = ora.step([a])
            steps += 1
            assert not err[0]
            assert np.array_equal(fov.view(np.uint32), f_o[0].view(np.uint32)) and np.array_equal(loc, l_o[0]), steps
            assert f32bits(gr) == gr_o.view(np.uint32)[0] and f32bits(orr) == lr_o.view(np.uint32)[0], steps
            assert (bool(gd), bool(ld)) == (bool(gd_o[0]), bool(ld_o[0])), steps
            if gd and rng.rand() < 0.7:
                reset_both()
                planned = False
                break
            if ld and rng.rand() < 0.8:
                break
    assert np.array_equal(np.asarray(env.state[2], np.float32).view(np.uint32), ora.export_visit()[0].view(np.uint32))
