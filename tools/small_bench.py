#!/usr/bin/env python
"""compact / incremental step kernels at scale (GPU box only): ms per step."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
def run(tag, n, variant, **kw):
    env = lmz.LmazeVecCuda(n, variant, seed=1, **kw)
    env.reset()
    a = torch.randint(0, 4, (4, n), device="cuda", dtype=torch.uint8)
    for i in range(5):
        env.step(a[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        env.step(a[i % 4])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 40
    print("%-16s n=%d  %.4f ms  %.2f G env-steps/s" % (tag, n, ms, n / ms / 1e6), flush=True)
    env.close()
for cap in ([int(c) for c in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]):
    t = (0, 0, cap, 0)
    print("CTAs per SM cap", cap)
    run("compact v0", 1 << 24, "v0", obs_mode="compact", tune=t)
    run("compact v3", 1 << 23, "v3", obs_mode="compact", tune=t)
    run("incr v0", 1 << 20, "v0", render_mode="incremental", tune=t)
    run("incr v3", 1 << 20, "v3", render_mode="incremental", tune=t)
    run("transition v0", 1 << 24, "v0", with_obs=False, tune=t)
