"""CPU tier: the product's own per-env FP32 arithmetic (urgym_env.cuh compiled for the host by tests/hostcheck)
against the FP64 oracle, all four tasks, capsule and hull geometry, with auto-reset through the shared
counter-based reset stream.  The CUDA kernels run exactly these functions, one env per thread (checked on the GPU
by test_gpu_parity.py)."""
import numpy as np
import pytest

from tests._hostcheck import HostCheckSim, ee_pose, philox
from tests._parity import run_parity, sta_moving_scenarios
from oracle import oracle_env as oe

TASKS = ["UR5OriReach-v1", "UR5ObsReach-v1", "UR5StaReach-v1", "UR5DynReach-v1"]
# hull geometry: FP32 link poses differ from the oracle's FP64 poses by up to ~1.5e-6 m, and so do the GJK distances
# built on them (pairs whose FP32 iteration does not converge are redone in FP64: urgym_device.cuh gjk_distance_refine).
# The budget: link distances 5e-6 m; rewards additionally 100 (Obs weight) x 5 links x 2 distances x 2e-6 m.
HULL_LD_TOL, HULL_REW_ATOL = 5e-6, 2e-3


def test_philox_matches_oracle_and_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert philox((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    rng = np.random.default_rng(0)
    for _ in range(50):
        c = tuple(int(x) for x in rng.integers(0, 2 ** 32, 4)); k = tuple(int(x) for x in rng.integers(0, 2 ** 32, 2))
        assert philox(c, k) == oe.philox4x32_10(c, k)


def test_fk_pose_tolerance():
    rng = np.random.default_rng(1)
    q = np.concatenate([[[0, -1.5708, 0, -1.5708, 0, 0]], rng.uniform(-2 * np.pi, 2 * np.pi, (2000, 6))]).astype(np.float32)
    ee = ee_pose(q)
    sim = oe.OracleSim()
    worst_p = worst_a = 0.0
    for i in range(len(q)):
        sim.set_joint_angles(q[i].astype(np.float64))
        p, e = sim.get_link_position(7), sim.get_link_orientation(7)
        if abs(abs(e[1]) - np.pi / 2) < 2e-2:       # Euler angles are ill-conditioned next to gimbal lock
            continue
        worst_p = max(worst_p, np.abs(ee[i, :3] - p).max())
        d = np.abs(ee[i, 3:] - e) % (2 * np.pi)
        worst_a = max(worst_a, np.minimum(d, 2 * np.pi - d).max())
    assert worst_p < 1e-5 and worst_a < 1e-5, (worst_p, worst_a)


@pytest.mark.parametrize("env_id", TASKS)
def test_capsule_rollout_parity(env_id):
    n, steps = 48, 70
    sim = HostCheckSim(env_id, oe.GEOM_CAPSULE, n, seed=1234)
    st = run_parity(sim, env_id, oe.GEOM_CAPSULE, n, steps, seed=1234)
    assert st["steps"] > 0.9 * n * steps and st["resets"] > 0, st


@pytest.mark.parametrize("env_id", TASKS)
def test_hull_rollout_parity(env_id):
    n, steps = 16, 40
    sim = HostCheckSim(env_id, oe.GEOM_HULL, n, seed=99, offset=1000)
    st = run_parity(sim, env_id, oe.GEOM_HULL, n, steps, seed=99, offset=1000, ld_tol=HULL_LD_TOL, rew_atol=HULL_REW_ATOL)
    assert st["steps"] > 0.9 * n * steps, st


@pytest.mark.parametrize("env_id", ["UR5ObsReach-v1", "UR5StaReach-v1", "UR5DynReach-v1"])
@pytest.mark.parametrize("geom", [oe.GEOM_CAPSULE, oe.GEOM_HULL])
def test_workbench_link_dist_mode_parity(env_id, geom):
    """link-distance mode "workbench" (per link min over obstacle, table, track: the docstring of
    pyb_setup.py:439-456 and the definition the shipped Obs / Sta policies need): observation
    columns, reset-time values and the reward term against the oracle in the same mode"""
    n, steps = (32, 50) if geom == oe.GEOM_CAPSULE else (10, 30)
    sim = HostCheckSim(env_id, geom, n, seed=21, offset=300, link_dist_mode=oe.LD_WORKBENCH)
    # hull geometry: FP32 GJK between a tessellated-cylinder hull and a box face parallel to it stalls 6e-5 short of the
    # optimum (upper arm above the track); this mode exists for the closed-loop check, a statistical comparison
    kw = dict(ld_tol=HULL_LD_TOL, rew_atol=HULL_REW_ATOL) if geom == oe.GEOM_HULL else {}
    st = run_parity(sim, env_id, geom, n, steps, seed=21, offset=300, link_dist_mode=oe.LD_WORKBENCH, **kw)
    assert st["steps"] > 0.85 * n * steps, st


@pytest.mark.parametrize("geom", [oe.GEOM_CAPSULE, oe.GEOM_HULL])
def test_sta_moving_obstacle_injection(geom):
    """ReachSta's 18-value set_goal_and_obstacle (reach.py:491-499) arms ReachSta.set_velocity (core.py:307-308,
    reach.py:518-541): the obstacle moves with twist (end - start) / 1 s until it is within 0.05 m of obstacle_end.
    obstacle_end / obstacle_start survive the resets that follow (reach.py:465-481 does not clear them), so the new
    episodes' obstacles drift with the stale twist -- replicated, and compared with the oracle step by step."""
    n, steps = (36, 60) if geom == oe.GEOM_CAPSULE else (8, 35)
    sc = sta_moving_scenarios(n, seed=5)

    def inject(sim, orc):
        sim.set_goal(sc[:, :6]); sim.set_obstacle(sc[:, 6:12]); sim.set_obstacle_start(sc[:, 6:12]); sim.set_obstacle_end(sc[:, 12:])
        sim.refresh()
        for i, e in enumerate(orc.envs):
            e.task.set_goal_and_obstacle(sc[i].astype(np.float64))

    sim = HostCheckSim("UR5StaReach-v1", geom, n, seed=8, offset=40)
    kw = dict(ld_tol=HULL_LD_TOL, rew_atol=HULL_REW_ATOL) if geom == oe.GEOM_HULL else {}
    st = run_parity(sim, "UR5StaReach-v1", geom, n, steps, seed=8, offset=40, action_scale=0.6, after_reset=inject, **kw)
    assert st["steps"] > 0.8 * n * steps and (st["resets"] > 0 or geom == oe.GEOM_HULL), st
    moved = np.linalg.norm(sim.E[:, 12:15] - sim.E[:, 6:9], axis=1)
    assert (moved > 0.05).sum() > n // 2                  # most scenarios really moved (end far from start)
