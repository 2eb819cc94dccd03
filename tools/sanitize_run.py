#!/usr/bin/env python3
"""Small workload for `compute-sanitizer --tool memcheck python tools/sanitize_run.py`: every kernel of the library on
ragged sizes, all four tasks, both geometries (hull: a few steps), chains, explicit resets, state access, host buffers."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import urgym_b200 as ug

g = torch.Generator(device="cuda").manual_seed(0)
for env_id in ("UR5OriReach-v1", "UR5ObsReach-v1", "UR5StaReach-v1", "UR5DynReach-v1"):
    for n in (200, 4099):
        env = ug.UR5VecEnv(env_id, n, seed=3, goal_buffers=True)
        env.reset()
        for t in range(30):
            env.step(torch.rand((n, 6), device="cuda", generator=g) * 3 - 1.5)
        env.reset(mask=(torch.rand(n, device="cuda", generator=g) < 0.3).to(torch.uint8))
        env.observe(); env.refresh()
        sd = env.state_dict(); env.load_state_dict(sd)
        if n > 1000:
            ring = [torch.rand((n, 6), device="cuda", generator=g) * 2 - 1 for _ in range(4)]
            gr = env.capture_steps(ring, chains=3)
            gr.replay(); gr.replay()
        buf = env.alloc_host_buffers()
        env.step_host(buf["actions"], buf)
        torch.cuda.synchronize()
        print(env_id, n, env.stats()["episodes"], flush=True)
        env.close()
env = ug.UR5VecEnv("UR5DynReach-v1", 300, seed=3, geometry="hull")
env.reset()
for t in range(3):
    env.step(torch.rand((300, 6), device="cuda", generator=g) * 2 - 1)
torch.cuda.synchronize()
print("hull ok", flush=True)
