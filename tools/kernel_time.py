#!/usr/bin/env python3
"""Step / auto-reset kernel durations of one build (URGYM_B200_LIB selects the library): CUDA events recorded by the
library around each kernel (urgym_profile_enable) over a queue of back-to-back steps.  Used for A/B runs while tuning.

    python tools/kernel_time.py [--task UR5DynReach-v1] [--envs 1048576] [--steps 64] [--geometry capsule]"""
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
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=120, help="untimed steps first: episode ages need ~100 steps to mix")
    ap.add_argument("--geometry", default="capsule")
    ap.add_argument("--tag", default=os.environ.get("URGYM_B200_LIB", "default"))
    a = ap.parse_args()
    import torch
    import urgym_b200 as ug
    env = ug.UR5VecEnv(a.task, a.envs, device=0, seed=0, geometry=a.geometry, goal_buffers=True)
    g = torch.Generator(device="cuda").manual_seed(1234)
    ring = [torch.rand((a.envs, 6), device="cuda", generator=g) * 2 - 1 for _ in range(8)]
    env.reset()
    for k in range(a.warmup):
        env.step(ring[k % 8])
    torch.cuda.synchronize()
    env.L.urgym_profile_enable(env.h, 1)
    for k in range(a.steps):
        env.step(ring[k % 8])
    s, r, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    env.L.urgym_profile_read(env.h, ctypes.byref(s), ctypes.byref(r), ctypes.byref(n))
    print(json.dumps({"tag": os.path.basename(a.tag), "task": a.task, "envs": a.envs, "step_kernel_us": 1e3 * s.value,
                      "reset_kernel_us": 1e3 * r.value, "steps": n.value}), flush=True)


if __name__ == "__main__":
    main()
