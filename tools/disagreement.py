#!/usr/bin/env python3
"""Capsule geometry (the throughput path) against hull geometry (the reference's geometry) at GPU scale: the same
states, the same actions, every step.

The hull simulator drives (auto-reset on, so the state distribution is the reference geometry's own steady state); before
every step its state is copied into a capsule simulator, both take the same actions, and the outcomes of that ONE step are
compared: the collision flag (hull = truth), the five link-obstacle distances, the reward.

    python tools/disagreement.py [--task UR5DynReach-v1] [--envs 1048576] [--steps 100] [--out profiles/disagreement_r02.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

LD_EDGES = [0.0, 1e-4, 3e-4, 1e-3, 3e-3, 1e-2, 3e-2, 1e-1, 1.0]


def measure(task="UR5DynReach-v1", envs=1 << 18, steps=100, warmup=100, seed=0, device=0):
    import torch
    import urgym_b200 as ug
    dev = torch.device("cuda", device)
    hull = ug.UR5VecEnv(task, envs, device=device, seed=seed, geometry="hull")
    cap = ug.UR5VecEnv(task, envs, device=device, seed=seed, geometry="capsule", auto_reset=False)
    hull.reset(); cap.reset()
    g = torch.Generator(device=dev).manual_seed(99)
    fields = ["goal", "q", "elapsed", "ep_return"]
    if task != "UR5OriReach-v1":
        fields = ["goal", "obstacle"] + (["obstacle_end"] if task == "UR5DynReach-v1" else []) + ["q", "elapsed", "ep_return", "link_dist"]
    for _ in range(warmup):                      # mixed episode ages
        hull.step(torch.rand((envs, 6), device=dev, generator=g) * 2 - 1)
    n = fp = fn = tp = 0
    succ_dis = 0
    ld_hist = torch.zeros(len(LD_EDGES) - 1, dtype=torch.int64, device=dev)
    ld_abs_sum = torch.zeros(5, dtype=torch.float64, device=dev)
    ld_near_n = 0
    ld_near_abs = torch.zeros(5, dtype=torch.float64, device=dev)
    rew_abs_sum, rew_n, rew_max = 0.0, 0, 0.0
    edges = torch.tensor(LD_EDGES, device=dev)
    for _ in range(steps):
        for f in fields:
            cap.set_state(f, hull.get_state(f))
        a = torch.rand((envs, 6), device=dev, generator=g) * 2 - 1
        _, rc, tc, _, ic = cap.step(a)
        ld_c = cap.get_state("link_dist") if task != "UR5OriReach-v1" else None
        rc, tc, sc = rc.clone(), tc.bool().clone(), ic["is_success"].bool().clone()
        # the hull simulator auto-resets: read its link distances of THIS step from the terminal / next observation rows
        _, rh, th, trh, ih = hull.step(a)
        th, sh = th.bool(), ih["is_success"].bool()
        coll_h, coll_c = th & ~sh, tc & ~sc
        n += envs
        tp += int((coll_h & coll_c).sum()); fp += int((~coll_h & coll_c).sum()); fn += int((coll_h & ~coll_c).sum())
        succ_dis += int((sh != sc).sum())
        both_free = ~coll_h & ~coll_c & ~sh & ~sc & ~trh.bool()
        if ld_c is not None:
            ld_h = hull.get_state("link_dist")              # envs that did not finish keep their link_dist of this step
            e = (ld_c - ld_h).abs()[both_free]
            ld_hist += torch.bincount(torch.bucketize(e.flatten(), edges[1:-1]), minlength=len(LD_EDGES) - 1)
            ld_abs_sum += e.double().sum(0)
            near = (ld_h < 0.2)[both_free]
            ld_near_abs += (e.double() * near).sum(0); ld_near_n += int(near.sum())
        dr = (rc - rh).abs()[both_free]
        rew_abs_sum += float(dr.double().sum()); rew_n += int(both_free.sum()); rew_max = max(rew_max, float(dr.max()) if dr.numel() else 0.0)
    res = {"task": task, "envs": envs, "steps": steps, "env_steps": n,
           "collision": {"hull_hits": tp + fn, "capsule_false_positives": fp, "capsule_false_negatives": fn,
                         "disagreement_rate_per_env_step": (fp + fn) / n,
                         "false_positive_share_of_hull_hits": fp / max(tp + fn, 1), "false_negative_share_of_hull_hits": fn / max(tp + fn, 1)},
           "success_flag_disagreements": succ_dis,
           "reward": {"mean_abs_error": rew_abs_sum / max(rew_n, 1), "max_abs_error": rew_max,
                      "over": "env-steps where neither geometry ended the episode"}}
    if task != "UR5OriReach-v1":
        res["link_dist"] = {"abs_error_histogram_edges_m": LD_EDGES, "abs_error_histogram": ld_hist.tolist(),
                            "mean_abs_error_per_link_m": (ld_abs_sum / max(rew_n, 1)).tolist(),
                            "mean_abs_error_per_link_when_hull_dist_below_0.2_m": (ld_near_abs / max(ld_near_n, 1) * 5).tolist(),
                            "note": "links 2..6 (upper arm, forearm, wrist 1-3) vs the obstacle; capsule minus hull"}
    hull.close(); cap.close()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="UR5DynReach-v1")
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    r = measure(a.task, a.envs, a.steps)
    print(json.dumps(r))
    if a.out:
        json.dump(r, open(a.out, "w"), indent=1)
