#!/usr/bin/env python
"""v3 TMA render: resident CTAs per SM x issuing warps (GPU box only)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
N = 1 << 20
for thr, cap in ((32, 1), (32, 2), (64, 1), (32, 0)):
    env = lmz.LmazeVecCuda(N, "v3", seed=1, tune=(thr, 0, cap, 0))
    env.reset()
    a = torch.randint(0, 4, (4, N), device="cuda", dtype=torch.uint8)
    for i in range(4):
        env.step(a[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        env.step(a[i % 4])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("v3 tma threads=%d ctas/sm<=%d: %.3f ms  %.1f M env-steps/s  %.0f GB/s" % (thr, cap, ms, N / ms / 1e3, N * 62222 / ms / 1e6), flush=True)
    env.close(); del env
