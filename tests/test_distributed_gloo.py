"""CPU tier, world_size 2 over gloo: the only multi-rank logic on the path is
(1) contiguous global-id sharding and (2) the integer all-reduce of episode
statistics at reporting time (never on the step path) -- SURVEY.md section 8e."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gym_lmaze_b200 import shard_range, allreduce_stats
    from gym_lmaze_b200._abi import STAT_NAMES
    from oracle import oracle as O
    lo, hi = shard_range(total, rank, world)
    # each rank advances its shard with the CPU oracle standing in for the device
    # (rank-local work has no collective); RNG is keyed by GLOBAL env id
    ora = O.OracleVec(O.V0, hi - lo, seed=3, env_id0=lo, autoreset=True)
    ora.reset(want_obs=False)
    gen = torch.Generator().manual_seed(100)
    acts = torch.randint(0, 4, (120, total), generator=gen)
    for t in range(120):
        ora.step(acts[t, lo:hi].numpy(), want_obs=False)
    local = dict(zip(STAT_NAMES, ora.stats.tolist()))
    summed = allreduce_stats(local, torch.device("cpu"))
    pos = ora.export()[0]
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, pos.tolist()))
    # bench.py's aggregation rule: value = sum of units / max over ranks of the time
    t = torch.tensor([10.0 + rank]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((summed, gathered, float(t)))
    dist.destroy_process_group()


def test_sharded_stats_allreduce_matches_single_process():
    total, world = 257, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    summed, gathered, tmax = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    from gym_lmaze_b200._abi import STAT_NAMES
    from oracle import oracle as O
    ora = O.OracleVec(O.V0, total, seed=3, env_id0=0, autoreset=True)
    ora.reset(want_obs=False)
    gen = torch.Generator().manual_seed(100)
    acts = torch.randint(0, 4, (120, total), generator=gen)
    for t in range(120):
        ora.step(acts[t].numpy(), want_obs=False)
    assert summed == dict(zip(STAT_NAMES, ora.stats.tolist()))
    pos = ora.export()[0].tolist()
    for lo, hi, p in gathered:
        assert p == pos[lo:hi]            # shard-invariant trajectories
    assert tmax == 11.0


def _hier_run(n, env_id0, lo_global, total, iters=40):
    """The planner / actor loop of lmaze-v5 on the CPU oracle for envs [lo_global, lo_global + n) of a batch of
    `total`: goals / actions are drawn for the WHOLE batch and sliced, spawns come from the global-id keyed RNG."""
    import numpy as np
    from oracle import oracle as O
    o = O.OracleHier(n, seed=5, env_id0=env_id0)
    o.reset(want_obs=False)
    rng = np.random.RandomState(9)
    mask = np.ones(n, np.uint8)
    n_gd = 0
    for _ in range(iters):
        goals = rng.randint(0, 25, size=total)[lo_global:lo_global + n]
        acts = rng.randint(0, 4, size=total)[lo_global:lo_global + n]
        o.planner_step(goals, mask=mask)
        _, _, _, _, gd, ld, _ = o.step(acts)
        if gd.any():
            o.reset(mask=gd, want_obs=False)
        mask = (gd | ld).astype(np.uint8)
        n_gd += int(gd.sum())
    return o.export().tolist(), n_gd


def _hier_worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gym_lmaze_b200 import shard_range
    lo, hi = shard_range(total, rank, world)
    state, n_gd = _hier_run(hi - lo, lo, lo, total)
    t = torch.tensor([n_gd], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                  # the only collective: integer statistics
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, state))
    if rank == 0:
        q.put((int(t), gathered))
    dist.destroy_process_group()


def test_hier_shards_reproduce_the_single_process_batch():
    """lmaze-v5: rank r of 2 owns a contiguous global-id range; trajectories (resets included, device-RNG spec keyed
    by global env id) are identical to the un-sharded batch and the summed episode count matches."""
    total, world = 131, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_hier_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    n_gd, gathered = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    state, n_gd_single = _hier_run(total, 0, 0, total)
    assert n_gd == n_gd_single and n_gd > 0
    for lo, hi, s in gathered:
        assert s == state[lo:hi]
