#!/usr/bin/env python
"""ST128 render path (lmz_env_st_kernel) CTA-size sweep (GPU box only)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
N = 1 << 19
for variant, nbytes in (("v0", 112910), ("v3", 62222)):
    for thr in (256, 512, 1024):
        env = lmz.LmazeVecCuda(N, variant, seed=1, render_mode="st128", tune=(thr, 0, 0, 0))
        env.reset()
        a = torch.randint(0, 4, (4, N), device="cuda", dtype=torch.uint8)
        for i in range(3):
            env.step(a[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            env.step(a[i % 4])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("st128 %s threads=%d: %.3f ms  %.1f M env-steps/s  %.0f GB/s" % (variant, thr, ms, N / ms / 1e3, N * nbytes / ms / 1e6), flush=True)
        env.close(); del env
