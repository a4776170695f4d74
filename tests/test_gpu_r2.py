"""GPU tier, round 2: regressions for the round-1 review (ADVICE.md / VERDICT.md) and the paths added in round 2.
Everything goes through the C ABI (gym_lmaze_b200.LmazeVecCuda / LmazeHierCuda) and is compared bit for bit with the
CPU oracle or with the golden fixtures produced by the unmodified reference."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RC_VALUE = [-0.0, -1.0, -0.01, 100.0]


@pytest.fixture(scope="module")
def lmz():
    import gym_lmaze_b200 as g
    from gym_lmaze_b200 import _abi
    _abi.load()                      # must be the in-tree CUDA library; raises if missing
    assert torch.cuda.is_available()
    return g


def rbits(t):
    return t.detach().cpu().numpy().view(np.uint32)


def unpack(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(np.float32)


def _rng_actions(O, seed, n, t):
    return np.array([O.rng_action(seed, i, t) for i in range(n)], dtype=np.int64)


# ---------------------------------------------------------------- incremental render + rollout
@pytest.mark.parametrize("variant", ["v0", "v3"])
def test_incremental_render_resyncs_after_rollout(lmz, oracle_mod, variant):
    """lmz_rollout moves every ball (and auto-resets) without rendering; the next step of an incremental-render
    env must fall back to a full render, or the tensor keeps the pre-rollout blocks (lmaze_env.py:208-234 is a
    full image every step)."""
    N, seed, T = 1500, 23, 37
    ov = oracle_mod.V0 if variant == "v0" else oracle_mod.V3
    ora = oracle_mod.OracleVec(ov, N, seed=seed, autoreset=True)
    env = lmz.LmazeVecCuda(N, variant, seed=seed, autoreset=True, render_mode="incremental")
    assert torch.equal(env.reset().cpu(), torch.from_numpy(ora.reset()))
    gen = torch.Generator().manual_seed(5)
    t_roll = 0
    for it in range(3):
        for _ in range(3):                                     # steps on the incremental kernel
            a = torch.randint(0, 5, (N,), generator=gen)
            o_ref, r_ref, d_ref = ora.step(a.numpy())
            obs, rew, done, _ = env.step(a)
            assert np.array_equal(rbits(rew), r_ref.view(np.uint32))
            assert torch.equal(obs.cpu(), torch.from_numpy(o_ref)), it
        rew, done = env.rollout(T)                             # no obs written
        for t in range(T):
            _, r_ref, d_ref = ora.step(_rng_actions(oracle_mod, seed, N, t_roll + t), want_obs=False)
            assert np.array_equal(rbits(rew[t]), r_ref.view(np.uint32)), (it, t)
        t_roll += T
    a = torch.randint(0, 4, (N,), generator=gen)
    o_ref, _, _ = ora.step(a.numpy())
    obs, _, _, _ = env.step(a)
    assert torch.equal(obs.cpu(), torch.from_numpy(o_ref))
    env.close()


def test_incremental_graph_replay_guard(lmz, oracle_mod):
    """A graph captured while the tensor was in sync replays the incremental kernel; after set_state / rollout the
    replay must refuse to run until the tensor is re-rendered."""
    N, seed = 600, 3
    ora = oracle_mod.OracleVec(oracle_mod.V0, N, seed=seed)
    env = lmz.LmazeVecCuda(N, "v0", seed=seed, render_mode="incremental")
    assert torch.equal(env.reset().cpu(), torch.from_numpy(ora.reset()))
    abuf = torch.zeros(N, dtype=torch.uint8, device="cuda")
    replay = env.capture_step(abuf)
    gen = torch.Generator().manual_seed(9)
    for _ in range(4):
        a = torch.randint(0, 4, (N,), generator=gen)
        abuf.copy_(a)
        replay()
        o_ref, _, _ = ora.step(a.numpy())
        assert torch.equal(env.obs.cpu(), torch.from_numpy(o_ref))
    env.set_state(env.get_state())                             # anything that desyncs the tensor
    with pytest.raises(RuntimeError):
        replay()
    env.render_obs()                                           # full render: back in sync
    a = torch.randint(0, 4, (N,), generator=gen)
    abuf.copy_(a)
    replay()
    o_ref, _, _ = ora.step(a.numpy())
    assert torch.equal(env.obs.cpu(), torch.from_numpy(o_ref))
    env.close()


# ---------------------------------------------------------------- test mode beats RANDOM_BALL = False
@pytest.mark.parametrize("random_ball,random_goal", [(False, True), (True, False), (False, False)])
def test_v3_test_mode_has_priority_over_fixed_flags(lmz, golden_dir, random_ball, random_goal):
    """lmaze_env_v3.py:145-146,154-155: mode == "test" is checked BEFORE RANDOM_GOAL / RANDOM_BALL."""
    z = np.load(os.path.join(golden_dir, "v3_table.npz"))
    env = lmz.LmazeVecCuda(7, "v3", random_ball=random_ball, random_goal=random_goal, seed=4)
    env.reset()                                                # a normal reset first (goal may have moved off 'X')
    obs = env.reset(mode="test")
    st = env.get_state().cpu().numpy()
    assert (st[:, 0:4] == np.array(z["test_info"])).all()      # ball (7,8), goal (8,8)
    assert np.array_equal(obs[3].cpu().numpy(), unpack(z["test_obs"], (3, 72, 72)))
    assert env.stats()["steps"] == 0                           # and no error was counted
    env.close()


def test_v3_test_mode_single_maze_class(lmz, golden_dir):
    z = np.load(os.path.join(golden_dir, "v3_table.npz"))
    env = lmz.LmazeEnv_v3(random_ball=False)
    obs = env.reset(mode="test")
    assert np.array_equal(obs, unpack(z["test_obs"], (3, 72, 72)))
    assert env.state_vector[0:4] == [int(v) for v in z["test_info"]]
    env.close()


# ---------------------------------------------------------------- set_state never repairs silently
def test_set_state_counts_unreachable_rows(lmz):
    env = lmz.LmazeVecCuda(6, "v0", autoreset=False)
    st = env.get_state()
    st[:, 0:2] = torch.tensor([[1, 1], [5, 5], [3, 3], [0, 4], [2, 2], [13, 1]], dtype=torch.int32)
    #                          'S' ok  'X' ok  'B' ok  border  wall    outside
    st[:, 5] = torch.tensor([0, 3, 2, 0, 0, 0], dtype=torch.int32)
    env.set_state(st)
    with pytest.raises(ValueError):
        env.stats()
    assert env.stats(check_errors=False)["steps"] == 0
    out = env.get_state().cpu().numpy()
    assert out[3, 0] == 1 and out[5, 0] == 10                  # clamped into the interior, and counted
    env.close()
    env = lmz.LmazeVecCuda(3, "v0", autoreset=False)
    st = env.get_state()
    st[:, 0:2] = torch.tensor([[1, 1], [5, 5], [10, 10]], dtype=torch.int32)
    env.set_state(st)
    env.stats()                                                # reachable rows: no error
    env.close()
    env = lmz.LmazeVecCuda(3, "v3", autoreset=False)
    st = env.get_state()
    st[:, 0:4] = torch.tensor([[4, 4, 8, 8], [4, 5, 0, 0], [4, 6, 8, 8]], dtype=torch.int32)   # row 1: goal on the border
    env.set_state(st)
    with pytest.raises(ValueError):
        env.stats()
    env.close()
    for variant in ("v2", "v4"):
        env = lmz.LmazeVecCuda(2, variant, autoreset=False)
        st = env.get_state()
        env.set_state(st)
        env.stats()
        st[1, 0] = 1                                           # the ball never leaves [2, 15]
        env.set_state(st)
        with pytest.raises(ValueError):
            env.stats()
        env.close()
    h = lmz.LmazeHierCuda(2, "v5", autoreset=False)
    h.reset()
    st = h.get_state()
    h.set_state(st)
    h.stats()
    st[0, 12] = 31                                             # foveal goal cell outside 0..24
    h.set_state(st)
    with pytest.raises(ValueError):
        h.stats()
    h.close()


# ---------------------------------------------------------------- bit-packed observations (72 B per v0 env)
@pytest.mark.parametrize("variant", ["v0", "v3"])
def test_bits_obs_reconstructs_reference_image(lmz, oracle_mod, variant):
    """obs_mode='bits': 1 bit per cell of the un-expanded layers; unpack + xE replication is the oracle's full image
    every step (auto-resets, invalid actions, a masked reset, a render window, the batch tail)."""
    N, T = 3001, 230
    ov = oracle_mod.V0 if variant == "v0" else oracle_mod.V3
    ora = oracle_mod.OracleVec(ov, N, seed=11, autoreset=True, threads=os.cpu_count() or 1)
    env = lmz.LmazeVecCuda(N, variant, seed=11, autoreset=True, obs_mode="bits")
    assert env.obs.dtype == torch.uint8 and tuple(env.obs.shape) == (N, 72 if variant == "v0" else 124)
    assert torch.equal(env.expand(env.reset()).cpu(), torch.from_numpy(ora.reset()))
    gen = torch.Generator().manual_seed(2)
    for t in range(T):
        a = torch.randint(0, 5, (N,), generator=gen)
        want = t % 5 == 0 or t > T - 3
        o_ref, r_ref, d_ref = ora.step(a.numpy(), want_obs=want)
        obs, rew, done, _ = env.step(a)
        assert np.array_equal(rbits(rew), r_ref.view(np.uint32)) and np.array_equal(done.cpu().numpy().view(np.uint8), d_ref), t
        if want:
            assert torch.equal(env.expand(obs).cpu(), torch.from_numpy(o_ref)), t
        if variant == "v3":
            assert int((obs[:, 121] >> 4).max()) == 0 and int(obs[:, 122:].max()) == 0     # pad bits stay zero
    # compact twin: the bits are exactly the compact bytes
    twin = lmz.LmazeVecCuda(N, variant, seed=11, obs_mode="compact")
    twin.set_state(env.get_state())
    assert torch.equal(env.unpack_bits(env.render_obs()), twin.render_obs())
    twin.close()
    # masked reset: only the masked rows change
    before = env.obs.clone()
    mask = torch.rand(N, generator=gen) < 0.25
    env.reset(mask=mask)
    assert torch.equal(env.obs[~mask.cuda()], before[~mask.cuda()])
    env.close()
    # render window
    env = lmz.LmazeVecCuda(N, variant, seed=11, obs_mode="bits", obs_window=1000)
    ora = oracle_mod.OracleVec(ov, N, seed=11)
    env.reset(); ora.reset(want_obs=False)
    for t in range(6):
        a = torch.randint(0, 4, (N,), generator=gen)
        o_ref, _, _ = ora.step(a.numpy())
        if t % 2 == 0:
            env.set_window(777 * (t // 2) + 1)
        obs, _, _, _ = env.step(a)
        assert torch.equal(env.expand(obs).cpu(), torch.from_numpy(o_ref[env.window_lo:env.window_lo + 1000])), t
    env.close()
    with pytest.raises(Exception):
        lmz.LmazeVecCuda(4, "v2", obs_mode="bits")


# ---------------------------------------------------------------- rollout with 1-byte reward codes
@pytest.mark.parametrize("variant", ["v0", "v3"])
def test_rollout_reward_codes(lmz, variant):
    N, T = 5000, 150
    a = lmz.LmazeVecCuda(N, variant, seed=5, with_obs=False)
    b = lmz.LmazeVecCuda(N, variant, seed=5, with_obs=False)
    a.reset(); b.reset()
    table = a.reward_table("cuda")
    assert rbits(table).tolist() == [0x80000000, 0xBF800000, 0xBC23D70A, 0x42C80000]
    for it in range(2):
        rew, done_a = a.rollout(T)
        codes, done_b = b.rollout(T, reward_codes=True)
        assert codes.dtype == torch.uint8 and int(codes.max()) <= 3
        assert torch.equal(table[codes.long()].view(torch.int32), rew.view(torch.int32))
        assert torch.equal(done_a, done_b)
    assert torch.equal(a.get_state(), b.get_state()) and a.stats() == b.stats()
    acts = torch.randint(0, 6, (T, N), device="cuda", dtype=torch.uint8)
    rew, _ = a.rollout(T, actions=acts)
    codes, _ = b.rollout(T, actions=acts, reward_codes=torch.empty((T, N), dtype=torch.uint8, device="cuda"))
    assert torch.equal(table[codes.long()].view(torch.int32), rew.view(torch.int32))
    a.close(); b.close()


# ---------------------------------------------------------------- double-buffered host pipeline
@pytest.mark.parametrize("variant,obs_mode", [("v0", "compact"), ("v0", "bits"), ("v3", "bits"), ("v0", "full")])
def test_host_pipeline_matches_oracle(lmz, oracle_mod, variant, obs_mode):
    """lmz_step_host_async / _wait: every output of every step (obs included) lands in pinned host memory one
    submit later and equals the oracle's; the full f32 observation stays on the device (with_obs=False)."""
    N, T = 4099, 40
    ov = oracle_mod.V0 if variant == "v0" else oracle_mod.V3
    ora = oracle_mod.OracleVec(ov, N, seed=21, autoreset=True)
    env = lmz.LmazeVecCuda(N, variant, seed=21, autoreset=True, obs_mode=obs_mode)
    env.reset(); ora.reset(want_obs=False)
    if obs_mode == "full":
        with pytest.raises(ValueError):
            env.host_pipeline()
    pipe = env.host_pipeline(with_obs=obs_mode != "full")
    gen = torch.Generator().manual_seed(4)
    acts = [torch.randint(0, 5, (N,), generator=gen, dtype=torch.uint8).pin_memory() for _ in range(T)]
    refs = []

    def check(res, ref):
        o_ref, r_ref, d_ref = ref
        assert np.array_equal(res.reward.numpy().view(np.uint32), r_ref.view(np.uint32))
        assert np.array_equal(res.done.numpy(), d_ref)
        if res.obs is not None:
            assert torch.equal(env.expand(res.obs.cuda()).cpu(), torch.from_numpy(o_ref))
    for t in range(T):
        refs.append(ora.step(acts[t].numpy().astype(np.int64)))
        res = pipe.submit(acts[t])
        assert (res is None) == (t == 0)
        if res is not None:
            check(res, refs[t - 1])
    check(pipe.drain(), refs[-1])
    assert pipe.drain() is None
    # the pipeline can be resumed after a drain, and the serial host call still works beside it
    refs.append(ora.step(acts[0].numpy().astype(np.int64)))
    assert pipe.submit(acts[0]) is None
    check(pipe.drain(), refs[-1])
    r = torch.empty(N, dtype=torch.float32).pin_memory(); d = torch.empty(N, dtype=torch.uint8).pin_memory()
    o_ref, r_ref, d_ref = ora.step(acts[1].numpy().astype(np.int64))
    env.step_host(acts[1], r, d)
    assert np.array_equal(r.numpy().view(np.uint32), r_ref.view(np.uint32)) and np.array_equal(d.numpy(), d_ref)
    assert torch.equal(env.expand(env.obs).cpu(), torch.from_numpy(o_ref))
    assert env.stats()["steps"] == (T + 2) * N
    env.close()
