"""Closed-loop statistical check with the policies the reference ships.

The reference publishes, for each task, the success rate and mean episode reward of its trained SAC policy over ~5 k
scenarios run in real PyBullet (model_test.py:26-61 with utils/generate.py:23-102; numbers in
Trained_Models/*/best.txt).  Those scenarios were random and not saved, so this is a *statistical* pin -- but it is
the only number in the reference that real PyBullet produced, and a policy trained on PyBullet observations only
succeeds here if frames, Euler conventions, link indexing, observation layout and obstacle geometry match.

Protocol (model_test.py:26-61): reset; inject the scenario; deterministic policy for at most 100 steps; an episode
ends at the first terminated step (or step 99); success = info["is_success"] there; reward summed up to there.
Scenarios (utils/generate.py): Ori -- goal positions on a 0.05 m grid x 5 random goal orientations;
Obs / Sta -- 5 000 natural resets; Dyn -- goal on the grid, obstacle start / end re-sampled until the target-to-end
distance is >= 0.1 m and start-end >= 0.3 m (reach.py:685-700)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHORT = {"UR5OriReach-v1": "Ori", "UR5ObsReach-v1": "Obs", "UR5StaReach-v1": "Sta", "UR5DynReach-v1": "Dyn"}


def load_policy(env_id, device):
    w = np.load(os.path.join(ROOT, "tests", "golden", f"policy_{SHORT[env_id]}.npz"))
    t = {k: torch.as_tensor(w[k], device=device) for k in w.files if not k.startswith("published")}
    pub = {k: float(w[k]) for k in w.files if k.startswith("published")}

    def act(obs):
        # SB3 CombinedExtractor: Dict space keys in sorted order -> achieved_goal, desired_goal, observation
        x = torch.cat([obs["achieved_goal"], obs["desired_goal"], obs["observation"]], dim=1)
        h = torch.relu(x @ t["latent_pi_0_weight"].T + t["latent_pi_0_bias"])
        h = torch.relu(h @ t["latent_pi_2_weight"].T + t["latent_pi_2_bias"])
        return torch.tanh(h @ t["mu_weight"].T + t["mu_bias"]).contiguous()
    return act, pub


def grid_positions(low, high, repeat=5):
    """utils/generate.py:30-47,66-86: int((high-low)/0.05)+1 points per axis (with the reference's float rounding)"""
    num = [int((high[i] - low[i]) / 0.05) + 1 for i in range(3)]
    pts = [[low[0] + i / 20, low[1] + j / 20, low[2] + k / 20]
           for i in range(num[0]) for j in range(num[1]) for k in range(num[2]) for _ in range(repeat)]
    return np.array(pts, np.float32)


def dyn_scenarios(goal_pos, seed=0):
    """reach.py:685-700 reset_generate: goal rotation, obstacle start and end re-sampled until accepted"""
    from oracle import oracle_env as oe
    rng = np.random.default_rng(seed)
    lo, hi = np.array([0.5, -0.8, 0.25]), np.array([1.2, 0.8, 0.75])
    sim = oe.OracleSim(geom=oe.GEOM_HULL)
    for nm, k in [("plane", "box"), ("table", "box"), ("track", "box"), ("target", "box"), ("obstacle", "cylinder")]:
        sim.add_body(nm, k)
    out = np.zeros((len(goal_pos), 18), np.float32)
    for n, gp in enumerate(goal_pos):
        while True:
            goal = np.concatenate([gp, oe.euler_constrained_from_uniform(rng.random(), rng.random())])
            start = np.concatenate([lo + (hi - lo) * rng.random(3), oe.euler_obstacle_from_uniform(*rng.random(3))])
            end = np.concatenate([lo + (hi - lo) * rng.random(3), oe.euler_obstacle_from_uniform(*rng.random(3))])
            sim.set_base_pose("target", goal[:3], goal[3:]); sim.set_base_pose("obstacle", end[:3], end[3:])
            if not (sim.get_target_to_obstacle_distance() < 0.1 or np.linalg.norm(end[:3] - start[:3]) < 0.3):
                break
        out[n] = np.concatenate([goal, start, end])
    return out


def run(env_id, geometry="capsule", seed=0, device=0, max_envs=None):
    import urgym_b200 as ug
    dev = torch.device("cuda", device)
    act, pub = load_policy(env_id, dev)
    if env_id in ("UR5OriReach-v1", "UR5DynReach-v1"):
        low = [0.3, -0.5, 0.0] if env_id == "UR5OriReach-v1" else [0.4, -0.5, 0.0]
        pos = grid_positions(np.array(low), np.array([0.75, 0.5, 0.2]))
        n = len(pos)
    else:
        n = 5000
    if max_envs:
        n = min(n, max_envs)
    env = ug.UR5VecEnv(env_id, n, device=device, seed=seed, geometry=geometry, auto_reset=False)
    env.reset()
    if env_id == "UR5OriReach-v1":
        goal = env.get_state("goal")
        goal[:, :3] = torch.as_tensor(pos[:n], device=dev)
        env.set_state("goal", goal)
    elif env_id == "UR5DynReach-v1":
        sc = torch.as_tensor(dyn_scenarios(pos[:n], seed), device=dev)
        env.set_state("goal", sc[:, :6].contiguous()); env.set_state("obstacle", sc[:, 6:12].contiguous())
        env.set_state("obstacle_end", sc[:, 12:].contiguous())
        env.refresh()
    obs = env.observe()
    finished = torch.zeros(n, dtype=torch.bool, device=dev)
    success = torch.zeros(n, dtype=torch.bool, device=dev)
    collided = torch.zeros(n, dtype=torch.bool, device=dev)
    ret = torch.zeros(n, dtype=torch.float64, device=dev)
    steps = torch.zeros(n, dtype=torch.int32, device=dev)
    for t in range(100):
        obs, rew, term, trunc, info = env.step(act(obs))
        live = ~finished
        ret += torch.where(live, rew.double(), torch.zeros_like(ret))
        ends = live & (term.bool() | (t == 99))
        success |= ends & info["is_success"].bool()
        collided |= ends & term.bool() & ~info["is_success"].bool()
        steps = torch.where(ends, torch.full_like(steps, t), steps)
        finished |= ends
    res = {"env_id": env_id, "geometry": geometry, "episodes": n,
           "success_rate_pct": 100.0 * float(success.double().mean()), "mean_reward": float(ret.mean()),
           "collision_rate_pct": 100.0 * float(collided.double().mean()), "mean_steps": float(steps.double().mean()),
           "published": pub}
    env.close()
    return res


if __name__ == "__main__":
    out = []
    for geometry in ("hull", "capsule"):
        for env_id in SHORT:
            r = run(env_id, geometry)
            out.append(r)
            print(json.dumps(r), flush=True)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "closed_loop.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)
