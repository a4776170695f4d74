#!/usr/bin/env python
"""lmaze-v5/v6 (planner/actor env) throughput at scale (GPU box only): the realistic loop
plannerStep(mask = localDone | globalDone) + step, with autoreset, timed with CUDA events.
Algorithmic bytes per env-step: 34,300 foveal + 19,600 local obs + 1,296 visit read (+1,296 written on the
steps whose local episode is over) + 12+12 state + 8 rewards + 4 flags + 1 action."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lmaze_b200 as lmz
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
threads = [int(t) for t in sys.argv[2].split(",")] if len(sys.argv) > 2 else [512]
split = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for thr in threads:
    env = lmz.LmazeHierCuda(N, "v5", seed=1, tune=(thr, 0, 0, split))
    env.reset()
    a = torch.randint(0, 4, (8, N), device="cuda", dtype=torch.uint8)
    g = torch.randint(0, 25, (8, N), device="cuda", dtype=torch.uint8)
    mask = torch.ones(N, dtype=torch.uint8, device="cuda")
    def loop(k, timing):
        global mask
        ts, tp, ld_frac = 0.0, 0.0, 0.0
        for i in range(k):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            env.plannerStep(g[i % 8], mask=mask)
            e[1].record()
            env.step(a[i % 8], goal_plane=False)
            e[2].record()
            mask = (env._ldone_u8 | env._done_u8)
            if timing:
                torch.cuda.synchronize()
                tp += e[0].elapsed_time(e[1]); ts += e[1].elapsed_time(e[2]); ld_frac += float(env._ldone_u8.float().mean())
        return tp / max(k, 1), ts / max(k, 1), ld_frac / max(k, 1)
    loop(12, False)
    torch.cuda.synchronize()
    tp, ts, ldf = loop(20, True)
    by = 34300 + 19600 + 1296 * (1 + ldf) + 24 + 8 + 4 + 1
    print("v5 N=%d threads=%d: step %.3f ms = %.1f M env-steps/s, %.0f GB/s (%.0f B/env-step, localDone %.2f); "
          "plannerStep(masked) %.3f ms; step+planner %.1f M env-steps/s"
          % (N, thr, ts, N / ts / 1e3, N * by / ts / 1e6, by, ldf, tp, N / (ts + tp) / 1e3), flush=True)
    print("stats", env.stats(check_errors=False))
    env.close(); del env
