"""Closed-loop statistical check with the policies the reference ships.

The reference publishes, for each task, the success rate and per-episode (reward, success, steps) lines of its trained
SAC policy over ~5 k scenarios run in real PyBullet (model_test.py:26-61 with utils/generate.py:23-102; numbers in
Trained_Models/*/best.txt).  Those scenarios were random and not saved, so this is a *statistical* pin -- but it is
the only output of real PyBullet the reference contains, and a policy trained on PyBullet observations only
succeeds here if frames, Euler conventions, link indexing, observation layout and scene geometry match.

Protocol (model_test.py:26-61): reset; inject the scenario; deterministic policy for at most 100 steps; an episode
ends at the first terminated step (or step 99); success = info["is_success"] there; reward summed up to there.
Scenarios (utils/generate.py): Ori -- goal positions on a 0.05 m grid x 5 random goal orientations;
Obs / Sta -- 5 000 natural resets; Dyn -- goal on the grid, obstacle start / end re-sampled until the target-to-end
distance is >= 0.1 m and start-end >= 0.3 m (reach.py:685-700).

LINK-DISTANCE MODE.  UR5ObsReach and UR5StaReach are run with link_dist="workbench" (per link the smallest distance to
obstacle, table and track).  Round 2's diagnosis (profiles/closed_loop_r02.json): the two policies
are quasi-linear in their five link_dist inputs; fed the distances the shipped get_link_distances computes (obstacle
only, pyb_setup.py:439-456) they dive wrist_1 into the table (Obs: 30 % collisions, all `wrist_1 - table`) or stall with
a constant action bias 5-8 cm beside the goal (Sta: 28 % time-outs); fed what the method's docstring describes ("distance
between workbench, obstacle and UR5 ... robot links to any obstacle") they reproduce the published success rate, time-out
share, collision share and step histogram (and, for Obs, the published reward sum, which contains the link_dist term).
Their zips carry start_time 2023-09-20/21; the UR5DynReach zip, 2024-06-07, works with the shipped definition."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHORT = {"UR5OriReach-v1": "Ori", "UR5ObsReach-v1": "Obs", "UR5StaReach-v1": "Sta", "UR5DynReach-v1": "Dyn"}
LD_COLS = {"UR5ObsReach-v1": (21, 26), "UR5StaReach-v1": (24, 29), "UR5DynReach-v1": (30, 35)}
STEP_BINS = [0, 3, 5, 7, 9, 12, 16, 25, 50, 100]
# the link-distance definition each shipped policy is evaluated with (see the module docstring)
POLICY_LINK_DIST = {"UR5OriReach-v1": "obstacle", "UR5ObsReach-v1": "workbench", "UR5StaReach-v1": "workbench",
                    "UR5DynReach-v1": "obstacle"}


def load_policy_numpy(env_id):
    w = np.load(os.path.join(ROOT, "tests", "golden", f"policy_{SHORT[env_id]}.npz"))
    t = {k: w[k].astype(np.float32) for k in w.files if not k.startswith("published")}
    pub = {k: (w[k].tolist() if w[k].ndim else float(w[k])) for k in w.files if k.startswith("published")}

    def act(ag, dg, obs):
        # SB3 CombinedExtractor: Dict space keys in sorted order -> achieved_goal, desired_goal, observation
        x = np.concatenate([ag, dg, obs], axis=1)
        h = np.maximum(x @ t["latent_pi_0_weight"].T + t["latent_pi_0_bias"], 0)
        h = np.maximum(h @ t["latent_pi_2_weight"].T + t["latent_pi_2_bias"], 0)
        return np.tanh(h @ t["mu_weight"].T + t["mu_bias"]).astype(np.float32)
    return act, pub


def load_policy(env_id, device):
    import torch
    w = np.load(os.path.join(ROOT, "tests", "golden", f"policy_{SHORT[env_id]}.npz"))
    t = {k: torch.as_tensor(w[k], device=device) for k in w.files if not k.startswith("published")}
    pub = {k: (w[k].tolist() if w[k].ndim else float(w[k])) for k in w.files if k.startswith("published")}

    def act(obs):
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


GRID_LOW = {"UR5OriReach-v1": [0.3, -0.5, 0.0], "UR5DynReach-v1": [0.4, -0.5, 0.0]}     # reach.py:151,584
GRID_HIGH = [0.75, 0.5, 0.2]


def axis_profiles(pos, success, steps):
    """The spatial profile of a grid run, as tests/golden/make_policy_fixtures.py extracts it from the published per-episode
    lines: per grid value of each goal-position axis, mean step count of the successes, failures, episodes."""
    pos, success, steps = np.asarray(pos, np.float64), np.asarray(success, bool), np.asarray(steps)
    out = {}
    for ax in range(3):
        cell = np.rint(pos[:, ax] * 20).astype(int)
        vals = np.unique(cell)
        out[f"axis{ax}_values"] = (vals / 20).tolist()
        out[f"axis{ax}_success_steps"] = [float(steps[success & (cell == v)].mean()) for v in vals]
        out[f"axis{ax}_success_steps_sd"] = [float(steps[success & (cell == v)].std()) for v in vals]
        out[f"axis{ax}_failures"] = [int((~success & (cell == v)).sum()) for v in vals]
        out[f"axis{ax}_episodes"] = [int((cell == v).sum()) for v in vals]
    return out


def compare_profiles(r, z_max=4.0, mean_step_tol=0.15):
    """The run's spatial profile against the published one.  Per grid value of every axis (250-1750 episodes each): the
    difference of the mean step counts of the successes in units of its standard error (sd of the step counts measured in
    this run, both sides), and the difference of the failure shares in units of the binomial standard error at the pooled
    share.  Asserts every |z| <= z_max, the mean absolute step difference over all grid values <= mean_step_tol, and a
    correlation >= 0.98 between the two x profiles (the steps a goal needs grow as it comes closer to the base: 5.4 -> 8.3).
    Returns (worst step z, worst failure z, mean |step difference|, correlation)."""
    pub, prof = r["published"], r["axis_profiles"]
    zs, zf, diffs = 0.0, 0.0, []
    for ax in range(3):
        assert np.allclose(pub[f"published_axis{ax}_values"], prof[f"axis{ax}_values"], atol=1e-6), (ax, prof[f"axis{ax}_values"])
        assert np.array_equal(pub[f"published_axis{ax}_episodes"], prof[f"axis{ax}_episodes"]), ax
        n = np.asarray(prof[f"axis{ax}_episodes"], float)
        a = np.asarray(pub[f"published_axis{ax}_success_steps"]); b = np.asarray(prof[f"axis{ax}_success_steps"])
        fa = np.asarray(pub[f"published_axis{ax}_failures"]) / n; fb = np.asarray(prof[f"axis{ax}_failures"]) / n
        se_s = np.asarray(prof[f"axis{ax}_success_steps_sd"]) * np.sqrt(2.0 / (n * (1.0 - fb)))
        pool = np.maximum(0.5 * (fa + fb), 0.01)
        se_f = np.sqrt(pool * (1.0 - pool) * 2.0 / n)
        zs = max(zs, float((np.abs(a - b) / se_s).max())); zf = max(zf, float((np.abs(fa - fb) / se_f).max()))
        diffs += np.abs(a - b).tolist()
    cx = float(np.corrcoef(pub["published_axis0_success_steps"], prof["axis0_success_steps"])[0, 1])
    res = (zs, zf, float(np.mean(diffs)), cx)
    assert zs <= z_max and zf <= z_max and res[2] <= mean_step_tol and cx >= 0.98, (res, prof)
    return res


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


def taxonomy(success, collided, steps, ret, pub, extra=None):
    """the figures the published per-episode lines allow to compare (tests/golden/make_policy_fixtures.py)"""
    success, collided = np.asarray(success, bool), np.asarray(collided, bool)
    steps, ret = np.asarray(steps), np.asarray(ret, np.float64)
    res = {"episodes": int(len(success)), "success_rate_pct": 100.0 * float(success.mean()),
           "mean_reward": float(ret.mean()), "collision_rate_pct": 100.0 * float(collided.mean()),
           "timeout_pct": 100.0 * float((~success & ~collided).mean()), "mean_steps": float(steps.mean()),
           "success_mean_steps": float(steps[success].mean()) if success.any() else None,
           "success_mean_reward": float(ret[success].mean()) if success.any() else None,
           "success_step_hist": np.histogram(steps[success], bins=STEP_BINS)[0].tolist(),
           "collision_mean_steps": float(steps[collided].mean()) if collided.any() else None,
           "published": pub}
    if extra:
        res.update(extra)
    return res


def run(env_id, geometry="capsule", seed=0, device=0, max_envs=None, link_dist=None, ld_ablation=None):
    """GPU run through the public batched API.  ld_ablation: None, "zero" (link_dist inputs of the POLICY set to 0) or
    "clip0.2" (min(link_dist, 0.2)): the policy-input ablations of the round-2 diagnosis; the simulator is untouched."""
    import torch
    import urgym_b200 as ug
    dev = torch.device("cuda", device)
    act, pub = load_policy(env_id, dev)
    link_dist = link_dist or POLICY_LINK_DIST[env_id]
    if env_id in ("UR5OriReach-v1", "UR5DynReach-v1"):
        pos = grid_positions(np.array(GRID_LOW[env_id]), np.array(GRID_HIGH))
        n = len(pos)
    else:
        n = 5000
    if max_envs:
        n = min(n, max_envs)
    env = ug.UR5VecEnv(env_id, n, device=device, seed=seed, geometry=geometry, auto_reset=False, link_dist=link_dist)
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
        pol_obs = obs
        if ld_ablation and env_id in LD_COLS:
            a, b = LD_COLS[env_id]
            o = obs["observation"].clone()
            o[:, a:b] = 0.0 if ld_ablation == "zero" else torch.clamp(o[:, a:b], max=0.2)
            pol_obs = {"observation": o, "achieved_goal": obs["achieved_goal"], "desired_goal": obs["desired_goal"]}
        obs, rew, term, trunc, info = env.step(act(pol_obs))
        live = ~finished
        ret += torch.where(live, rew.double(), torch.zeros_like(ret))
        ends = live & (term.bool() | (t == 99))
        success |= ends & info["is_success"].bool()
        collided |= ends & term.bool() & ~info["is_success"].bool()
        steps = torch.where(ends, torch.full_like(steps, t), steps)
        finished |= ends
    extra = {"env_id": env_id, "geometry": geometry, "link_dist": link_dist, "policy_ld_ablation": ld_ablation}
    if env_id in GRID_LOW and n == len(pos):
        extra["axis_profiles"] = axis_profiles(pos, success.cpu().numpy(), steps.cpu().numpy())
    res = taxonomy(success.cpu().numpy(), collided.cpu().numpy(), steps.cpu().numpy(), ret.cpu().numpy(), pub, extra)
    env.close()
    return res


def run_host(env_id, geom=1, n=500, seed=0, link_dist_mode=0, ld_ablation=None, grid=False):
    """The same protocol on tests/hostcheck (the product's per-env code compiled for the host): the CPU tier's
    closed-loop check, natural-reset scenarios for every task; grid=True (UR5OriReach, UR5DynReach): the reference's own scenario
    grid of goal positions (utils/generate.py:23-47,66-86), all of it, with the spatial profile in the result."""
    from tests._hostcheck import HostCheckSim, GOAL_DIM, TASK_ID
    act, pub = load_policy_numpy(env_id)
    pos = None
    if grid:
        assert env_id in GRID_LOW
        pos = grid_positions(np.array(GRID_LOW[env_id]), np.array(GRID_HIGH))
        n = len(pos)
    sim = HostCheckSim(env_id, geom, n, seed=seed, autoreset=False, link_dist_mode=link_dist_mode)
    obs = sim.reset()
    G = GOAL_DIM[TASK_ID[env_id]]
    if grid and env_id == "UR5OriReach-v1":
        goal = obs[:, 12:12 + G].copy()
        goal[:, :3] = pos
        sim.set_goal(goal)
        obs = sim.observe()
    elif grid:
        sc = dyn_scenarios(pos, seed)
        sim.set_goal(sc[:, :6]); sim.set_obstacle(sc[:, 6:12]); sim.set_obstacle_end(sc[:, 12:])
        sim.refresh()                           # tail of set_goal_and_obstacle (reach.py:702-713): link_dist, collision
        obs = sim.observe()
    finished = np.zeros(n, bool); success = np.zeros(n, bool); collided = np.zeros(n, bool)
    steps = np.zeros(n, int); ret = np.zeros(n)
    for t in range(100):
        o = obs.copy()
        if ld_ablation and env_id in LD_COLS:
            a, b = LD_COLS[env_id]
            o[:, a:b] = 0.0 if ld_ablation == "zero" else np.minimum(o[:, a:b], 0.2)
        out = sim.step(act(o[:, :G], o[:, 12:12 + G], o))
        obs = out["obs"]
        live = ~finished
        ret += np.where(live, out["reward"], 0)
        ends = live & (out["terminated"] | (t == 99))
        success |= ends & out["is_success"]
        collided |= ends & out["collision"]
        steps[ends] = t
        finished |= ends
        if finished.all():
            break
    extra = {"env_id": env_id, "geometry": "capsule" if geom == 1 else "hull",
             "link_dist": "workbench" if link_dist_mode else "obstacle", "policy_ld_ablation": ld_ablation,
             "scenarios": "goal-position grid (host instantiation)" if grid else "natural resets (host instantiation)"}
    if grid:
        extra["axis_profiles"] = axis_profiles(pos, success, steps)
    return taxonomy(success, collided, steps, ret, pub, extra)


if __name__ == "__main__":
    out = []
    for geometry in ("hull", "capsule"):
        for env_id in SHORT:
            r = run(env_id, geometry)
            out.append(r)
            print(json.dumps({k: v for k, v in r.items() if k != "published"}), flush=True)
    # the diagnosis: the shipped definition and the two policy-input ablations, hull geometry
    for env_id in ("UR5ObsReach-v1", "UR5StaReach-v1", "UR5DynReach-v1"):
        for kw in (dict(link_dist="obstacle"), dict(link_dist="obstacle", ld_ablation="zero"),
                   dict(link_dist="obstacle", ld_ablation="clip0.2"), dict(link_dist="workbench")):
            if env_id != "UR5DynReach-v1" and kw == dict(link_dist="workbench"):
                continue        # already in the first block
            r = run(env_id, "hull", **kw)
            out.append(r)
            print(json.dumps({k: v for k, v in r.items() if k != "published"}), flush=True)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "closed_loop.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)
