#!/usr/bin/env python
"""One launch sequence over every kernel of the package at a representative size (for ncu; GPU box only).

    ncu --set full --clock-control none --profile-from-start off -k regex:lmz_ -o X python tools/profile_final.py MANIFEST.json

Writes MANIFEST.json: the hash of the kernel sources this build was compiled from (gym_lmaze_b200.build.source_hash)
and the ordered list of `lmz_*` launches (tag, expected kernel name fragment, envs, algorithmic bytes per env-step),
which tools/ncu_summary.py joins with ncu's raw page by launch order."""
import json, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
from gym_lmaze_b200.build import source_hash

launches = []
prof = torch.cuda.profiler          # ncu --profile-from-start off: only launches between start() and stop() are captured
prof.start()


def unprofiled(fn, n):
    """n warm-up calls outside the capture.  60 steps put the visit histories of v4 / v5 near their steady-state length
    (warp maximum ~45 of ~50); the episodes are fully de-synchronised only after ~1,000 steps (bench.py warms that long),
    but under ncu even an uncaptured launch costs ~0.1 s: a 1,000-step warm-up per variant ran past a 13-minute limit."""
    torch.cuda.synchronize(); prof.stop()
    for i in range(n):
        fn(i)
    torch.cuda.synchronize(); prof.start()


def note(tag, kernel, envs, bytes_per_env=None, key=None):
    launches.append({"tag": tag, "kernel": kernel, "envs": envs, "bytes_per_env_step": bytes_per_env, "traffic_key": key})


def run(tag, n, variant, kernel, bytes_per_env, key=None, steps=1, warm=0, **kw):
    env = lmz.LmazeVecCuda(n, variant, seed=1, **kw)
    env.reset()
    note(tag + " reset", kernel if kw.get("render_mode") != "incremental" else "lmz_env_tma_kernel", n)
    a = torch.randint(0, env.num_actions, (4, n), device="cuda", dtype=torch.uint8)
    if warm:
        unprofiled(lambda i: env.step(a[i % 4]), warm)
    for i in range(steps):
        env.step(a[i % 4])
        first_incr = kw.get("render_mode") == "incremental" and i == 0
        note(tag + " step", kernel, n, bytes_per_env, key if i == steps - 1 else None)
    torch.cuda.synchronize()
    return env


run("v0 full render (TMA)", 1 << 19, "v0", "lmz_env_tma_kernel", 112910, "v0_tma").close()
run("v0 full render (ST128)", 1 << 18, "v0", "lmz_env_st_kernel", 112910, "v0_st128", render_mode="st128").close()
run("v3 full render (TMA)", 1 << 19, "v3", "lmz_env_tma_kernel", 62222, "v3_tma").close()
run("v2 full render", 1 << 20, "v2", "lmz_env_fov_kernel", 24514, "v2_tma").close()
run("v4 full render, steady state (60 steps in)", 1 << 19, "v4", "lmz_env_fov_kernel", 34379, "v4_tma", steps=2, warm=60).close()
n = 1 << 19
h = lmz.LmazeHierCuda(n, "v5", seed=1)
h.reset(); note("v5 reset", "lmz_env_fov_kernel", n)
a = torch.randint(0, 4, (4, n), device="cuda", dtype=torch.uint8)
g = torch.randint(0, 25, (4, n), device="cuda", dtype=torch.uint8)
def hier_step(i):
    h.plannerStep(g[i % 4], mask="auto"); h.step(a[i % 4], goal_plane=False)
h.plannerStep(g[0], mask="auto"); note("v5 plannerStep (auto mask), first: every env", "lmz_planner_kernel", n)
h.step(a[0], goal_plane=False); note("v5 step, first", "lmz_env_fov_kernel", n, 54020)
unprofiled(hier_step, 60)
h.plannerStep(g[1], mask="auto"); note("v5 plannerStep (auto mask), steady state", "lmz_planner_kernel", n)
h.step(a[1], goal_plane=False); note("v5 step, steady state (61 steps in)", "lmz_env_fov_kernel", n, 54020, "v5_tma")
rew = h.rollout(32); note("v5 rollout T=32 (planner + actor)", "lmz_fov_rollout_kernel", n)
torch.cuda.synchronize(); h.close()
run("v0 compact u8", 1 << 22, "v0", "lmz_env_compact_kernel", 590, obs_mode="compact").close()
run("v0 bit-packed", 1 << 22, "v0", "lmz_env_compact_kernel", 86, obs_mode="bits").close()
run("v3 compact u8", 1 << 22, "v3", "lmz_env_compact_kernel", 986, obs_mode="compact").close()
e = lmz.LmazeVecCuda(1 << 19, "v0", seed=1, render_mode="incremental")
e.reset(); note("v0 incremental: full render (reset)", "lmz_env_tma_kernel", 1 << 19)
a = torch.randint(0, 4, (4, 1 << 19), device="cuda", dtype=torch.uint8)
for i in range(2):
    e.step(a[i]); note("v0 incremental step #%d" % i, "lmz_env_incr_kernel", 1 << 19, 406)
torch.cuda.synchronize(); e.close()
run("v2 compact f32 crops", 1 << 22, "v2", "lmz_fov_small_kernel", 514, obs_mode="compact").close()
run("v4 compact f32 crops, steady state (60 steps in)", 1 << 21, "v4", "lmz_fov_small_kernel", 779, obs_mode="compact", steps=2, warm=60).close()
n = 1 << 21
h = lmz.LmazeHierCuda(n, "v5", seed=1, obs_mode="compact")
h.reset(); note("v5 compact reset", "lmz_fov_small_kernel", n)
a = torch.randint(0, 4, (4, n), device="cuda", dtype=torch.uint8)
g = torch.randint(0, 25, (4, n), device="cuda", dtype=torch.uint8)
unprofiled(hier_step, 60)
h.plannerStep(g[1], mask="auto"); note("v5 compact plannerStep (auto mask), steady state", "lmz_fov_small_kernel", n)
h.step(a[1], goal_plane=False); note("v5 compact step, steady state (61 steps in)", "lmz_fov_small_kernel", n, 1100 + 64 + 37)
torch.cuda.synchronize(); h.close()
e = lmz.LmazeVecCuda(1 << 22, "v0", seed=1, with_obs=False)
e.reset(); note("v0 transition-only reset", "lmz_env_compact_kernel", 1 << 22)
a = torch.randint(0, 4, (1 << 22,), device="cuda", dtype=torch.uint8)
e.step(a); note("v0 transition-only step", "lmz_env_compact_kernel", 1 << 22, 14)
e.rollout(64); note("v0 rollout T=64", "lmz_rollout_kernel", 1 << 22, 5.25 * 64)
e.rollout(64, reward_codes=True); note("v0 rollout T=64, u8 reward codes", "lmz_rollout_kernel", 1 << 22, 2.25 * 64)
torch.cuda.synchronize(); e.close()
for v in ("v2", "v4"):
    e = lmz.LmazeVecCuda(1 << 21, v, seed=1, with_obs=False)
    e.reset(); note(v + " transition-only reset", "lmz_fov_small_kernel", 1 << 21)
    e.rollout(64); note(v + " rollout T=64", "lmz_fov_rollout_kernel", 1 << 21)
    torch.cuda.synchronize(); e.close()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_final_manifest.json"
json.dump({"csrc_hash": source_hash(), "launches": launches}, open(out, "w"), indent=1)
torch.cuda.synchronize(); prof.stop()
print("done: %d lmz launches" % len(launches), flush=True)
