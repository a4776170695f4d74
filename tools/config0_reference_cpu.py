#!/usr/bin/env python
"""BASELINE configs[0]: one lmaze-v0 env, 10,000-step random-action rollout through the UNMODIFIED reference's
step() on CPU (build container only -- the reference does not travel to the GPU box).

    python tools/config0_reference_cpu.py [steps] [processes]

random.seed(0) drives the reference's own spawn draws, actions come from random.Random(1).randrange(4), reset() on
done.  Prints one JSON line: steps/s of one process and the aggregate of P independent ones.
"""
import contextlib
import io
import json
import multiprocessing as mp
import os
import random
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def rollout(steps):
    from oracle.ref_loader import load_reference_module
    mod = load_reference_module("v0")
    random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        env = mod.LmazeEnv()
    acts = random.Random(1)
    env.reset()
    ret, episodes = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        obs, r, d, info = env.step(acts.randrange(4))
        ret += r
        if d:
            env.reset()
            episodes += 1
    dt = time.perf_counter() - t0
    return steps / dt, ret, episodes, tuple(obs.shape), str(obs.dtype)


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    one = rollout(steps)
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(rollout, [steps // 4] * procs)
    wall = time.perf_counter() - t0
    print(json.dumps({
        "config": "BASELINE configs[0]: lmaze_env_v0 single env, %d-step random-action rollout via the reference step() on CPU" % steps,
        "impl": "unmodified reference (gym_lmaze/envs/lmaze_env.py) under tests/_gymstub",
        "steps_per_s_one_process": one[0], "return": one[1], "episodes": one[2], "obs": [list(one[3]), one[4]],
        "processes": procs, "steps_per_s_aggregate": sum(r[0] for r in res),
        "host": "build container (not the GPU box): %d logical CPUs" % (os.cpu_count() or 1), "python": sys.version.split()[0],
    }))
