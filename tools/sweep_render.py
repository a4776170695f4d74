#!/usr/bin/env python
"""Sweep the launch-tuning knobs of the fused step kernel at full scale (GPU box only).

    python tools/sweep_render.py [--envs N] [--variant v0] > gpurun_out/sweep.txt

Prints one line per configuration: average kernel time over K steps (CUDA events)
and the algorithmic HBM write bandwidth, next to pure-write (fill) and copy
baselines measured in the same process, so the render kernel can be judged
against what the memory system gives a trivial streaming kernel on this very GPU.
"""
import argparse
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz  # noqa: E402


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--variant", default="v0")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    N = args.envs
    step_bytes = (112896 if args.variant == "v0" else 62208) + 14
    results = []

    # ---- baselines on a buffer of the same size as the obs tensor
    nbytes = N * (step_bytes - 14)
    buf = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
    for name, fn, moved in (
            ("torch fill_(1.0)  [pure write]", lambda: buf.fill_(1.0), nbytes),
            ("torch zero_()     [pure write]", lambda: buf.zero_(), nbytes),
            ("torch copy half->half [read+write]", lambda: buf[: buf.numel() // 2].copy_(buf[buf.numel() // 2:]), nbytes)):
        fn()
        ms = timed(fn, 5)
        print("%-44s %8.3f ms  %8.1f GB/s" % (name, ms, moved / ms / 1e6), flush=True)
        results.append({"name": name, "ms": ms, "gbs": moved / ms / 1e6})
    del buf
    torch.cuda.empty_cache()

    gen = torch.Generator(device="cuda").manual_seed(1)
    ring = torch.randint(0, 4, (4, N), generator=gen, device="cuda", dtype=torch.uint8)

    def run(render_mode, tune):
        env = lmz.LmazeVecCuda(N, args.variant, seed=1, render_mode=render_mode, tune=tune)
        env.reset()
        for i in range(3):
            env.step(ring[i % 4])
        k = [0]

        def one():
            env.step(ring[k[0] % 4]); k[0] += 1
        ms = timed(one, args.steps)
        env.close()
        del env
        return ms

    combos = [("st128", (0, 0, 0, 0)), ("st128", (512, 0, 0, 0)), ("st128", (256, 0, 0, 0))]
    if args.quick:
        combos += [("tma", (0, 0, 0, 0)), ("tma", (64, 0, 0, 0))]
    else:
        for thr in (32, 64, 128, 256):
            combos.append(("tma", (thr, 0, 0, 0)))
        for pol in (1, 2):
            combos.append(("tma", (32, pol, 0, 0)))
        for split in (4704, 37632):
            combos.append(("tma", (32, 0, 0, split)))
    for mode, tune in combos:
        ms = run(mode, tune)
        gbs = N * step_bytes / ms / 1e6
        print("%-6s threads=%-4d l2pol=%d order=%d split=%-6d  %8.3f ms  %8.1f GB/s  %6.2f M env-steps/s"
              % (mode, tune[0], tune[1], tune[2], tune[3], ms, gbs, N / ms / 1e3), flush=True)
        results.append({"mode": mode, "tune": tune, "ms": ms, "gbs": gbs})
    print(json.dumps(results))


if __name__ == "__main__":
    main()
