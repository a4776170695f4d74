#!/usr/bin/env python
"""(threads per CTA) x (CTAs per SM) sweep of the foveal kernel (GPU box only)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
def run(variant, N, nbytes, thr, cap):
    hier = variant == "v5"
    env = (lmz.LmazeHierCuda(N, "v5", seed=1, tune=(thr, 0, cap, 0)) if hier
           else lmz.LmazeVecCuda(N, variant, seed=1, tune=(thr, 0, cap, 0)))
    env.reset()
    na = 4 if hier else 25
    a = torch.randint(0, na, (4, N), device="cuda", dtype=torch.uint8)
    g = torch.randint(0, 25, (4, N), device="cuda", dtype=torch.uint8)
    def step(i):
        if hier:
            env.plannerStep(g[i % 4], mask="auto"); env.step(a[i % 4], goal_plane=False)
        else:
            env.step(a[i % 4])
    for i in range(int(os.environ.get("WARM", "4"))):      # WARM=60: episodes in steady state (visit histories at their usual length)
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12):
        step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 12
    print("%s threads=%4d ctas/sm<=%d: %.3f ms  %.1f M env-steps/s  %.0f GB/s" % (variant, thr, cap, ms, N / ms / 1e3, N * nbytes / ms / 1e6), flush=True)
    env.close(); del env
cfgs = [(int(c.split("x")[0]), int(c.split("x")[1])) for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else "224x1,256x1,512x1,1024x1".split(","))]
sizes = {"v2": (1 << 22, 24514), "v4": (1 << 21, 36906), "v5": (1 << 20, 56400)}
for variant in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["v2", "v4", "v5"]):
    for thr, cap in cfgs:
        run(variant, sizes[variant][0], sizes[variant][1], thr, cap)
