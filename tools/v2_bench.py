#!/usr/bin/env python
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
N = 1 << 22
for rep in range(2):
  for thr in (512, 1024):
    env = lmz.LmazeVecCuda(N, "v2", seed=1, tune=(thr, 0, 0, 0))
    env.reset()
    a = torch.randint(0, 25, (4, N), device="cuda", dtype=torch.uint8)
    for i in range(5):
        env.step(a[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        env.step(a[i % 4])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("v2 threads=%d: %.3f ms  %.1f M env-steps/s  %.0f GB/s" % (thr, ms, N / ms / 1e3, N * 24514 / ms / 1e6), flush=True)
    env.close(); del env
