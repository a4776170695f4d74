"""Time the UNMODIFIED reference's own Python step() loop on the host cores (bench.py's cpu_baseline leg only).

BASELINE.md section 4: P independent processes, each `env = LmazeEnv()`; random actions; `env.reset()` on done;
a fixed wall budget per process.  The reference is looked for in baseline/_ref (the offline `pip install --target`
of /root/reference, which travels to the GPU box) and then in /root/reference (build container).  Test
infrastructure: nothing in gym_lmaze_b200/ imports this file.
"""
import contextlib
import io
import multiprocessing as mp
import os
import random
import sys
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CANDIDATE_ROOTS = (os.path.join(_ROOT, "baseline", "_ref"), "/root/reference")

_N_ACTIONS = {"v0": 4, "v2": 25, "v3": 4, "v4": 25}
_V3_WORDS = ("left", "right", "up", "down")          # lmaze_env_v3.py:236-247 compares strings


def find_reference_root():
    for root in CANDIDATE_ROOTS:
        if os.path.isfile(os.path.join(root, "gym_lmaze", "envs", "lmaze_env.py")):
            return root
    return None


def _worker(args):
    variant, root, budget_s, k = args
    os.environ["LMAZE_REFERENCE_ROOT"] = root
    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)
    from oracle import ref_loader
    ref_loader.REFERENCE_ROOT = root
    mod = ref_loader.load_reference_module(variant)
    random.seed(k)
    acts = random.Random(1000 + k)
    with contextlib.redirect_stdout(io.StringIO()):
        env = getattr(mod, ref_loader._FILES[variant][1])()
        env.reset()
        n_act = _N_ACTIONS[variant]
        steps = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < budget_s or steps < 3:
            a = acts.randrange(n_act)
            out = env.step(_V3_WORDS[a] if variant == "v3" else a)
            if out[2]:
                env.reset()
            steps += 1
        dt = time.perf_counter() - t0
    return steps, dt


def measure(variant, procs, budget_s):
    """{"value": aggregate env-steps/s over `procs` processes, ...} or None when the reference is not present
    (or the variant has no single-call step loop: v5's planner/actor protocol is timed through the port only)."""
    root = find_reference_root()
    if root is None or variant not in _N_ACTIONS:
        return None
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_worker, [(variant, root, budget_s, k) for k in range(procs)])
    per_proc = [s / dt for s, dt in res]
    return {"value": sum(per_proc), "unit": "env-steps/s", "cores": procs, "kind": "reference",
            "per_process": sum(per_proc) / len(per_proc), "steps": sum(s for s, _ in res),
            "sample": "the unmodified reference's step() loop (%s, loaded from %s under tests/_gymstub), %d independent "
                      "processes x %.1f s, random actions, reset() on done" % (ref_loader_file(variant), root, procs, budget_s)}


def ref_loader_file(variant):
    return {"v0": "gym_lmaze/envs/lmaze_env.py", "v2": "gym_lmaze/envs/lmaze_env_v2.py",
            "v3": "gym_lmaze/envs/lmaze_env_v3.py", "v4": "gym_lmaze/envs/lmaze_env_v4.py"}[variant]
