#!/usr/bin/env python
"""CTAs-per-SM sweep of the foveal compact kernel lmz_fov_small_kernel (GPU box only): ms per step.
    python tools/fov_compact_sweep.py [variants] [caps] [threads]      e.g.  v2,v4,v5  0,1,2,3  64,128,256"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
def run(variant, N, cap, thr=0):
    hier = variant == "v5"
    env = (lmz.LmazeHierCuda(N, "v5", seed=1, obs_mode="compact", tune=(thr, 0, cap, 0)) if hier
           else lmz.LmazeVecCuda(N, variant, seed=1, obs_mode="compact", tune=(thr, 0, cap, 0)))
    env.reset()
    a = torch.randint(0, 4 if hier else 25, (4, N), device="cuda", dtype=torch.uint8)
    g = torch.randint(0, 25, (4, N), device="cuda", dtype=torch.uint8)
    def step(i):
        if hier:
            env.plannerStep(g[i % 4], mask="auto"); env.step(a[i % 4], goal_plane=False)
        else:
            env.step(a[i % 4])
    for i in range(int(os.environ.get("WARM", "5"))):      # WARM=60: episodes in steady state
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print("%s compact N=%d threads=%d ctas/sm<=%d: %.4f ms  %.2f G env-steps/s" % (variant, N, thr, cap, ms, N / ms / 1e6), flush=True)
    env.close(); del env
sizes = {"v2": 1 << 24, "v4": 1 << 23, "v5": 1 << 23}
caps = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
thrs = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
for variant in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["v2", "v4", "v5"]):
    for thr in thrs:
        for cap in caps:
            run(variant, sizes[variant], cap, thr)
