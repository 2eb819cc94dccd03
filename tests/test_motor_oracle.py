"""CPU tier: the motor-path oracle (oracle/ur_motor_oracle.c, PARITY UNPINNED against PyBullet) checked against independent
computations of the same mechanics: the dynamics must be a consistent Lagrangian system whatever Bullet does around it."""
import numpy as np
import pytest

from oracle import motor_oracle as mo
from oracle import oracle_env as oe

RNG = np.random.default_rng(7)


def _rand_q():
    return RNG.uniform(-2.5, 2.5, 6)


def test_mass_matrix_symmetric_positive_definite():
    for _ in range(20):
        M = mo.mass_matrix(_rand_q())
        assert np.abs(M - M.T).max() < 1e-12
        assert np.linalg.eigvalsh(M).min() > 1e-3


def test_gravity_torque_is_the_gradient_of_the_potential():
    for _ in range(10):
        q = _rand_q()
        g = mo.bias(q, np.zeros(6), damping=0.0)            # bias at rest = gravity torque = dV/dq
        num = np.zeros(6)
        for k in range(6):
            dq = np.zeros(6); dq[k] = 1e-6
            num[k] = (mo.energy(q + dq, np.zeros(6)) - mo.energy(q - dq, np.zeros(6))) / 2e-6
        assert np.abs(g - num).max() < 1e-6


def test_kinetic_energy_matches_link_velocities():
    # T = 1/2 qd^T M qd must equal the sum over bodies computed from finite-difference link motion of the ee (mass 1,
    # inertia 1): checked through the ee velocity Jacobian instead -- v_ee = d(pos)/dt along qd
    for _ in range(10):
        q, qd = _rand_q(), RNG.uniform(-1, 1, 6)
        _, v = mo.ee_state(q, qd)
        p1, _ = mo.ee_state(q + 1e-6 * qd, qd)
        p0, _ = mo.ee_state(q - 1e-6 * qd, qd)
        assert np.abs(v - (p1 - p0) / 2e-6).max() < 1e-6


def test_free_motion_conserves_energy_to_first_order_in_dt():
    q0, qd0 = np.array([0.3, -1.0, 0.5, 0.2, -0.4, 0.1]), np.array([0.5, -0.3, 0.2, 1.0, -1.0, 0.5])
    e0 = mo.energy(q0, qd0)
    drift = []
    for dt in (1e-3, 1e-4):
        q, qd = mo.substeps(q0, qd0, q0, force=None, n_sub=int(round(0.1 / dt)), dt=dt, damping=0.0)
        drift.append(abs(mo.energy(q, qd) - e0))
    assert drift[1] < 0.2 * drift[0] and drift[1] < 5e-3          # semi-implicit Euler: O(dt)
    # link damping only ever removes energy
    q, qd = mo.substeps(q0, qd0, q0, force=None, n_sub=100, dt=1e-3)
    q2, qd2 = mo.substeps(q0, qd0, q0, force=None, n_sub=100, dt=1e-3, damping=0.0)
    assert mo.energy(q, qd) < mo.energy(q2, qd2)


def test_motors_hold_the_reset_pose_and_respect_the_torque_limit():
    q, qd = mo.substeps(mo.NEUTRAL, np.zeros(6), mo.NEUTRAL)
    assert np.abs(q - mo.NEUTRAL).max() < 1e-5 and np.abs(qd).max() < 1e-4
    # a full-scale action: the shoulder cannot gain more velocity per substep than force * dt / (its smallest inertia)
    target = mo.NEUTRAL + 0.1 * np.pi
    q1, qd1 = mo.substeps(mo.NEUTRAL, np.zeros(6), target, n_sub=1)
    Minv = np.linalg.inv(mo.mass_matrix(mo.NEUTRAL))
    bound = np.abs(Minv) @ (mo.JOINT_FORCES * mo.DT) + mo.DT * np.abs(Minv @ mo.bias(mo.NEUTRAL, np.zeros(6)))
    assert np.all(np.abs(qd1) <= bound + 1e-9)
    # with unlimited force one substep reaches kp * error / dt exactly (the rows decouple at the fixed point)
    q2, qd2 = mo.substeps(mo.NEUTRAL, np.zeros(6), target, force=np.full(6, 1e9), n_sub=1, iters=400)
    assert np.abs(qd2 - mo.KP * (target - mo.NEUTRAL) / mo.DT).max() < 1e-6


def test_env_protocol():
    env = mo.UR5IAIReachOracle(oe.PhiloxStream(3), env_index=5)
    o = env._get_obs()
    assert o["observation"].shape == (6,) and o["observation"].dtype == np.float32
    assert np.all(env.goal >= mo.GOAL_LOW) and np.all(env.goal <= mo.GOAL_HIGH)
    assert np.abs(o["observation"][3:]).max() == 0.0
    r_prev = None
    for k in range(100):
        o, r, term, trunc, info = env.step(np.zeros(6))
        assert r == pytest.approx(-float(np.linalg.norm(o["achieved_goal"] - o["desired_goal"])), abs=1e-6)
        assert not term
        r_prev = r
    assert trunc and env.elapsed == 100 and r_prev is not None
