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
