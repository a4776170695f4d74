#!/usr/bin/env python
"""Per-step times of a long run of a compact kernel next to nvidia-smi clocks / power (GPU box only): burst vs sustained.
    python tools/sustained_probe.py v2|v4|v0 [steps]"""
import os, subprocess, sys, threading, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
variant = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 800
N = {"v0": 1 << 24, "v2": 1 << 24, "v4": 1 << 23}[variant]
env = lmz.LmazeVecCuda(N, variant, seed=1, obs_mode="compact")
env.reset()
a = torch.randint(0, env.num_actions, (4, N), device="cuda", dtype=torch.uint8)
for i in range(64):
    env.step(a[i % 4])
torch.cuda.synchronize()
samples, stop = [], False
def pump():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-i", "0"],
                             stdout=subprocess.PIPE, text=True).stdout.strip()
        samples.append((time.perf_counter(), out))
        time.sleep(0.05)
th = threading.Thread(target=pump); th.start()
t0 = time.perf_counter()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
evs[0].record()
for i in range(steps):
    env.step(a[i % 4]); evs[i + 1].record()
torch.cuda.synchronize()
stop = True; th.join()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
for lo in range(0, steps, steps // 10):
    seg = sorted(ms[lo:lo + steps // 10])
    print("%s steps %4d-%4d: min %.3f median %.3f max %.3f ms" % (variant, lo, lo + steps // 10, seg[0], seg[len(seg) // 2], seg[-1]))
for t, s in samples[:: max(1, len(samples) // 12)]:
    print("t=%.2fs  sm,mem MHz, W, C, reasons: %s" % (t - t0, s))
