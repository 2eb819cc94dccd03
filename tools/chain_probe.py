#!/usr/bin/env python3
"""Where the whole step's time goes beyond the step kernel: ms per env step of CUDA-graph replays for a given number of
chains, with the auto-reset on and off (off: finished envs simply keep stepping, so only the step kernels run), next to
the eager per-kernel event timings of the same state.

    python tools/chain_probe.py [--chains 1 2 4] [--envs 1048576]"""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="UR5DynReach-v1")
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--chains", type=int, nargs="+", default=[1, 2, 4])
    ap.add_argument("--graph-steps", type=int, default=32)
    ap.add_argument("--replays", type=int, default=6)
    a = ap.parse_args()
    import torch
    import urgym_b200 as ug
    from urgym_b200 import _native as nat
    env = ug.UR5VecEnv(a.task, a.envs, device=0, seed=0, geometry="capsule", goal_buffers=True)
    g = torch.Generator(device="cuda").manual_seed(1234)
    ring = [torch.rand((a.envs, 6), device="cuda", generator=g) * 2 - 1 for _ in range(8)]
    env.reset()
    for k in range(150):
        env.step(ring[k % 8])
    for autoreset in (1, 0):
        nat.check(env.h, env.L.urgym_set_autoreset(env.h, autoreset))
        for c in a.chains:
            graph = env.capture_steps(ring * (a.graph_steps // 8), chains=c)
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.replays):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (a.replays * a.graph_steps)
            env.L.urgym_profile_enable(env.h, 1)
            for k in range(32):
                env.step(ring[k % 8])
            s, r, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
            env.L.urgym_profile_read(env.h, ctypes.byref(s), ctypes.byref(r), ctypes.byref(n))
            env.L.urgym_profile_enable(env.h, 0)
            print(json.dumps({"autoreset": autoreset, "chains": c, "graph_us_per_step": 1e3 * ms,
                              "eager_step_kernel_us": 1e3 * s.value, "eager_reset_kernel_us": 1e3 * r.value}), flush=True)
        if autoreset:        # bring the batch back to a mixed-age state for the second pass
            pass


if __name__ == "__main__":
    main()
