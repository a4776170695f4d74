"""CPU tier: the arithmetic of the device's visit FOLD (lmz_fov.cuh, visit_fold2), restated in numpy float32, gives the
oracle's float visit layer bit for bit.

The CUDA kernels keep lmaze-v4's `state[2]` (lmaze_env_v4.py:116-119,211-214: `(state[2] + visitMap) / 2` over all 324
cells, float64 expression stored as float32) as the HISTORY of window centres and evaluate a cell as the weighted sum
    u <- RN32(u + m_k * 2^(k-T)),  k = 0 .. T-1        (m_k = 1 iff the window of averaging k covers the cell)
one fma per entry.  This test drives the C oracle -- which performs the reference's literal full-layer averaging and is
pinned to the reference's outputs (tests/test_oracle_golden.py) -- through episodes without a reset, rebuilds every
cell from the ball positions with exactly that recurrence, and compares bit patterns.  (The GPU tier compares the
kernels' layers with the oracle's; this test is the same statement about the FORMULA, runnable without a GPU.)"""
import numpy as np
import pytest


def _fold(centres, T):
    """all 324 cells after the first T averagings; float32 arithmetic, one rounding per entry (an fma with an exact product)"""
    u = np.zeros((18, 18), dtype=np.float32)
    xs, ys = np.meshgrid(np.arange(18), np.arange(18), indexing="ij")
    for k in range(T):
        cx, cy = centres[k]
        m = ((np.abs(xs - cx) <= 2) & (np.abs(ys - cy) <= 2)).astype(np.float64)
        w = np.float64(2.0) ** (k - T)                                   # exact: a power of two
        u = (u.astype(np.float64) + m * w).astype(np.float32)            # exact sum in float64, ONE rounding to float32
    return u


def _halving(centres, T):
    """the reference's own expression, float64 then stored as float32 (lmaze_env_v4.py:116-119)"""
    v = np.zeros((18, 18), dtype=np.float32)
    xs, ys = np.meshgrid(np.arange(18), np.arange(18), indexing="ij")
    for k in range(T):
        cx, cy = centres[k]
        m = ((np.abs(xs - cx) <= 2) & (np.abs(ys - cy) <= 2)).astype(np.float64)
        v = ((v.astype(np.float64) + m) / 2).astype(np.float32)
    return v


def test_weighted_sum_equals_the_reference_expression_on_random_histories():
    rng = np.random.RandomState(3)
    for trial in range(300):
        T = int(rng.randint(1, 65))
        if trial % 3 == 0:                       # a ball that hardly moves: cells covered 25+ times in a row (RN to 1.0 and ties)
            c = np.clip(np.cumsum(rng.randint(-1, 2, size=(T, 2)), axis=0) + 8, 2, 15)
        else:                                    # teleporting fovea: sparse covers, long gaps (sticky low bits)
            c = rng.randint(2, 16, size=(T, 2))
        a, b = _fold(c, T), _halving(c, T)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (trial, T)


@pytest.mark.parametrize("seed", [1, 2])
def test_fold_of_the_ball_positions_is_the_oracles_layer(seed):
    from oracle import oracle as O
    n, steps = 48, 63                                            # reset averages once (v4), + 63 steps = 64 entries
    ora = O.OracleVec(O.V4, n, seed=seed, autoreset=False)
    ora.reset(want_obs=False)
    rng = np.random.RandomState(seed)
    centres = [[tuple(p[:2])] for p in ora.export()[0]]           # the spawn cell: lmaze_env_v4.py:110-119
    for t in range(steps):
        ora.step(rng.randint(0, 25, size=n), want_obs=False)      # stepping on past `done` is legal with autoreset off
        for i, p in enumerate(ora.export()[0]):
            centres[i].append(tuple(p[:2]))
        if t in (0, 5, 24, 40, steps - 1):
            layer = ora.export_visit()
            for i in range(n):
                mine = _fold(centres[i], len(centres[i]))
                assert np.array_equal(mine.view(np.uint32), layer[i].view(np.uint32)), (t, i)
