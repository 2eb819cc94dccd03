#!/usr/bin/env python3
"""Calibrate the capsule geometry of the throughput path against the reference's hull geometry.

TEST / BUILD-TIME TOOLING (lives under oracle/): uses the oracle's hull-geometry distances as the measuring device
and writes ur-gym_b200/csrc/urgym_capsule_fit.h (committed; read by the product and by the oracle's capsule mode).

For every pair class of PyBullet.check_collision / get_link_distances (pyb_setup.py:382-456) the capsule geometry
computes  d = dist(segment cores) - R.  Bounding capsules (R = bounding radius) are exact-safe but fat: with the
reference's trained policies 22 % of the episodes ended on a forearm-wrist_2 "collision" the hulls do not have.  Here
R is chosen per pair class so that the collision boolean (d <= 0.01) disagrees least with the hull geometry on
configurations the tasks visit (random-action rollouts and rollouts of the shipped policies), ties broken towards the
median distance offset near contact.  The obstacle's axis-segment half length is searched as well.

    python oracle/calibrate_capsules.py [n_episodes]     -> prints the fit and the residual disagreement, writes the header
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_env as oe  # noqa: E402

PAIRS = [(1, 3), (1, 4), (1, 5), (1, 6), (2, 4), (2, 5), (2, 6), (3, 5), (3, 6)]
THR = 0.01


def policy(short):
    w = np.load(os.path.join(ROOT, "tests", "golden", f"policy_{short}.npz"))

    def act(o):
        x = np.concatenate([o["achieved_goal"], o["desired_goal"], o["observation"]])
        h = np.maximum(w["latent_pi_0_weight"] @ x + w["latent_pi_0_bias"], 0)
        h = np.maximum(w["latent_pi_2_weight"] @ h + w["latent_pi_2_bias"], 0)
        return np.tanh(w["mu_weight"] @ h + w["mu_bias"]).astype(np.float32)
    return act


def collect(n_episodes, obst_h, seed=0):
    """(hull distances [n,24], core distances [n,24]) over visited configurations; episodes run in HULL geometry"""
    rng = np.random.default_rng(seed)
    z7, z9 = (ctypes.c_double * 7)(), (ctypes.c_double * 9)()
    oe.lib().orc_set_capsule_fit(z7, z7, z9, ctypes.c_double(obst_h))
    hull, core = [], []
    plans = [("UR5StaReach-v1", None), ("UR5DynReach-v1", None), ("UR5ObsReach-v1", None),
             ("UR5OriReach-v1", "Ori"), ("UR5DynReach-v1", "Dyn"), ("UR5StaReach-v1", "Sta"), ("UR5ObsReach-v1", "Obs")]
    for env_id, pol in plans:
        env = oe.make(env_id, geom=oe.GEOM_HULL, stream=oe.NumpyStream(seed + 1))
        act = policy(pol) if pol else None
        for ep in range(n_episodes):
            o, _ = env.reset()
            for t in range(60):
                a = act(o) if act else rng.uniform(-1, 1, 6).astype(np.float32)
                o, r, term, trunc, info = env.step(a)
                h = env.sim.all_pair_distances()
                c = env.sim.capsule_core_distances()
                if env_id == "UR5OriReach-v1":
                    h[:5] = np.nan; c[:5] = np.nan
                hull.append(h); core.append(c)
                if term:
                    break
    return np.array(hull), np.array(core)


def best_margin(hull, core):
    """R minimising the number of (core - R <= THR) != (hull <= THR); ties -> closest to the median offset near contact"""
    ok = np.isfinite(hull) & np.isfinite(core)
    hull, core = hull[ok], core[ok]
    near = hull < 0.08
    if near.sum() < 20:
        near = hull < np.partition(hull, min(len(hull) - 1, 50))[min(len(hull) - 1, 50)]
    med = float(np.median((core - hull)[near]))
    cands = np.linspace(med - 0.03, med + 0.03, 241)
    truth = hull <= THR
    dis = np.array([np.count_nonzero(((core - R) <= THR) != truth) for R in cands])
    best = cands[dis == dis.min()]
    R = float(best[np.argmin(np.abs(best - med))])
    pred = (core - R) <= THR
    return R, dict(n=int(len(hull)), hull_hits=int(truth.sum()), false_pos=int((pred & ~truth).sum()), false_neg=int((~pred & truth).sum()),
                   median_offset=med, abs_err_near=float(np.mean(np.abs(core - R - hull)[near])))


def main():
    n_ep = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    results = {}
    for obst_h in (0.15, 0.165, 0.18, 0.2):
        hull, core = collect(n_ep, obst_h)
        tot = 0
        fit = []
        for l in range(2, 7):
            R, st = best_margin(hull[:, l - 2], core[:, l - 2])
            fit.append((R, st)); tot += st["false_pos"] + st["false_neg"]
        results[obst_h] = (tot, fit, hull, core)
        print(f"obstacle half length {obst_h}: obstacle-pair disagreements {tot}")
    obst_h = min(results, key=lambda k: results[k][0])
    _, fit_o, hull, core = results[obst_h]
    R_obst, R_box, R_self = np.zeros(7), np.zeros(7), np.zeros(9)
    report = []
    for l in range(2, 7):
        R_obst[l] = fit_o[l - 2][0]
        report.append((f"link {l} vs obstacle", fit_o[l - 2]))
        hb = np.concatenate([hull[:, 5 + l - 2], hull[:, 10 + l - 2]]); cb = np.concatenate([core[:, 5 + l - 2], core[:, 10 + l - 2]])
        R, st = best_margin(hb, cb)
        R_box[l] = R - 0.001                      # the box keeps its own margin of 0.001
        report.append((f"link {l} vs table/track", (R, st)))
    for p, (a, b) in enumerate(PAIRS):
        R, st = best_margin(hull[:, 15 + p], core[:, 15 + p])
        R_self[p] = R
        report.append((f"self pair {a}:{b}", (R, st)))
    M = np.load(os.path.join(ROOT, "ur-gym_b200", "assets", "ur5e_model.npz"))
    r_b = M["capsule_r"] + 0.001
    print(f"\nchosen obstacle half length {obst_h}")
    tp = tfp = tfn = 0
    for name, (R, st) in report:
        print(f"  {name:24s} R = {R:.4f}   n {st['n']:6d}  hull hits {st['hull_hits']:5d}  false+ {st['false_pos']:4d}  false- {st['false_neg']:4d}"
              f"  |err| near contact {1e3 * st['abs_err_near']:.1f} mm")
        tp += st["hull_hits"]; tfp += st["false_pos"]; tfn += st["false_neg"]
    print(f"  total: hull hits {tp}, false positives {tfp}, false negatives {tfn}")
    note = (f"CALIBRATED by oracle/calibrate_capsules.py on {len(hull)} configurations ({n_ep} episodes per plan: random-action and "
            f"shipped-policy rollouts):\n * per pair class, the margin minimising disagreement of the collision boolean with the hull geometry;\n"
            f" * residual on the calibration set: {tfp} false positives, {tfn} false negatives against {tp} hull hits.")
    hdr = open(os.path.join(ROOT, "ur-gym_b200", "csrc", "urgym_capsule_fit.h")).read()
    head = hdr[:hdr.index(" * ", hdr.index("p: (1:3"))]
    body = (head + " * " + note + "\n */\n#ifndef URGYM_CAPSULE_FIT_H\n#define URGYM_CAPSULE_FIT_H\n"
            f"static const double URGYM_FIT_OBST_H = {obst_h!r};\n"
            "static const double URGYM_FIT_OBST[7] = {" + ", ".join(repr(float(x)) for x in R_obst) + "};\n"
            "static const double URGYM_FIT_BOX[7] = {" + ", ".join(repr(float(x)) for x in R_box) + "};\n"
            "static const double URGYM_FIT_SELF[9] = {" + ", ".join(repr(float(x)) for x in R_self) + "};\n#endif\n")
    open(os.path.join(ROOT, "ur-gym_b200", "csrc", "urgym_capsule_fit.h"), "w").write(body)
    print("wrote ur-gym_b200/csrc/urgym_capsule_fit.h;  bounding values were: obstacle", np.round(r_b[2:] + 0.05, 4), "box", np.round(r_b[2:], 4))


if __name__ == "__main__":
    main()
