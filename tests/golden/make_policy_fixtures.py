#!/usr/bin/env python3
"""Extract the ACTOR weights of the reference's shipped SAC policies (data, not source) and the closed-loop numbers the
reference publishes for them, as fixtures for the closed-loop check.

Run in the build container (where /root/reference is mounted); outputs are committed because the GPU box has no
reference:   python tests/golden/make_policy_fixtures.py  ->  tests/golden/policy_{Ori,Obs,Sta,Dyn}.npz

Sources (reference file): Trained_Models/Trained_*/best_model.zip:policy.pth (SB3 SAC MultiInputPolicy actor:
features = concat(achieved_goal, desired_goal, observation) -> Linear 256 -> ReLU -> Linear 256 -> ReLU -> mu -> tanh),
Trained_Models/Trained_{Ori,Obs,Sta}/best.txt:1-2 and Trained_Models/Trained_Dyn/best_modeltest_result.txt:1-2
(success rate and mean episode reward over the model_test.py scenarios)."""
import io
import json
import os
import re
import sys
import zipfile

import numpy as np
import torch

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
STEP_BINS = [0, 3, 5, 7, 9, 12, 16, 25, 50, 100]
GRID_LOW = {"Ori": [0.3, -0.5, 0.0], "Dyn": [0.4, -0.5, 0.0]}      # ReachOri / ReachDyn goal_range_low (reach.py:154,601)
RESULT = {"Ori": "best.txt", "Obs": "best.txt", "Sta": "best.txt", "Dyn": "best_modeltest_result.txt"}

for t in ("Ori", "Obs", "Sta", "Dyn"):
    d = os.path.join(REF, "Trained_Models", f"Trained_{t}")
    z = zipfile.ZipFile(os.path.join(d, "best_model.zip"))
    sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=False)
    out = {k.replace("actor.", "").replace(".", "_"): v.numpy().astype(np.float32) for k, v in sd.items()
           if k.startswith("actor.") and "log_std" not in k}
    lines = open(os.path.join(d, RESULT[t])).read().splitlines()
    out["published_success_rate_pct"] = np.float64(re.search(r"([-\d.]+)%", lines[0]).group(1))
    out["published_mean_reward"] = np.float64(re.search(r"is ([-\d.]+)", lines[1]).group(1))
    eps = np.array([[float(x) for x in l.replace(" ", "").split(",")] for l in lines[2:] if l.strip()])
    out["published_episodes"] = np.int64(len(eps))
    out["published_mean_steps"] = np.float64(eps[:, 2].mean())
    # failure taxonomy of the published per-episode lines (reward, success, last step index): model_test.py:40-60
    rew, ok, st = eps[:, 0], eps[:, 1] == 1, eps[:, 2]
    out["published_timeout_pct"] = np.float64(100.0 * ((st == 99) & ~ok).mean())
    out["published_collision_pct"] = np.float64(100.0 * ((st < 99) & ~ok).mean())
    out["published_success_mean_steps"] = np.float64(st[ok].mean())
    out["published_success_mean_reward"] = np.float64(rew[ok].mean())
    out["published_success_step_hist"] = np.histogram(st[ok], bins=STEP_BINS)[0].astype(np.int64)
    out["published_collision_mean_steps"] = np.float64(st[(st < 99) & ~ok].mean())
    # Spatial profile of the published per-episode lines.  For Ori and Dyn the scenario list is a grid of goal positions in a
    # fixed order (utils/generate.py:30-47,66-86: 0.05 m steps, x outermost, 5 scenarios per grid point), so line n of the
    # result file belongs to a known goal position: per grid value of each axis, the mean step count of the successful
    # episodes, the number of failures and the number of episodes.
    if t in GRID_LOW:
        low, high = GRID_LOW[t], [0.75, 0.5, 0.2]
        num = [int((high[i] - low[i]) / 0.05) + 1 for i in range(3)]
        idx = np.array([[i, j, k] for i in range(num[0]) for j in range(num[1]) for k in range(num[2]) for _ in range(5)])
        assert len(idx) == len(eps), (t, len(idx), len(eps))
        for ax in range(3):
            out[f"published_axis{ax}_values"] = np.array([low[ax] + v / 20 for v in range(num[ax])], np.float64)
            out[f"published_axis{ax}_success_steps"] = np.array([st[ok & (idx[:, ax] == v)].mean() for v in range(num[ax])])
            out[f"published_axis{ax}_failures"] = np.array([(~ok & (idx[:, ax] == v)).sum() for v in range(num[ax])], np.int64)
            out[f"published_axis{ax}_episodes"] = np.array([(idx[:, ax] == v).sum() for v in range(num[ax])], np.int64)
    # when the policy was trained (SB3 `start_time`, ns since the epoch) and with which library versions
    meta = json.loads(z.read("data"))
    out["published_train_start_ns"] = np.int64(meta["start_time"])
    out["published_num_timesteps"] = np.int64(meta["num_timesteps"])
    np.savez_compressed(os.path.join(HERE, f"policy_{t}.npz"), **out)
    print(t, {k: (v.shape if getattr(v, "shape", ()) else float(v)) for k, v in out.items()})
