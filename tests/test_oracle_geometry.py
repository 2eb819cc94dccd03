"""Tier T2 self-checks: the oracle's FK and GJK distances against independent numpy/scipy computations.

These do not pin the oracle to PyBullet (impossible here, see oracle/ur_oracle_sim.c header); they show that
the restated *geometry* is computed correctly: distance(A, B) = max over unit n of the separation along n
(support-function duality), maximised by scipy on the sphere, must equal the GJK answer."""
import ctypes
import os

import numpy as np
import pytest
from scipy.optimize import minimize

from oracle import oracle_env as oe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL = np.load(os.path.join(ROOT, "ur-gym_b200", "assets", "ur5e_model.npz"))


def rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return (np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]) @ np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
            @ np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]]))


def fk_numpy(q):
    """Independent FK with 4x4 homogeneous transforms straight from the URDF numbers."""
    T = np.eye(4)
    out = [T.copy()]
    for i in range(6):
        J = np.eye(4)
        J[:3, :3] = rpy(*MODEL["joint_rpy"][i]); J[:3, 3] = MODEL["joint_xyz"][i]
        Rq = np.eye(4); Rq[:3, :3] = rpy(0, 0, q[i])
        T = T @ J @ Rq
        out.append(T.copy())
    return out


def oracle_fk(q):
    qd = (ctypes.c_double * 6)(*q)
    pos = ((ctypes.c_double * 3) * 7)(); rot = ((ctypes.c_double * 9) * 7)()
    oe.lib().orc_fk(qd, pos, rot)
    return np.array([list(p) for p in pos]), np.array([list(r) for r in rot]).reshape(7, 3, 3)


def test_fk_matches_independent_numpy_and_survey_values():
    rng = np.random.default_rng(0)
    for q in [np.zeros(6), np.array([0, -1.5708, 0, -1.5708, 0, 0])] + list(rng.uniform(-3, 3, (50, 6))):
        pos, rot = oracle_fk(q)
        T = fk_numpy(q)
        for l in range(7):
            np.testing.assert_allclose(pos[l], T[l][:3, 3], atol=1e-13)
            np.testing.assert_allclose(rot[l], T[l][:3, :3], atol=1e-13)
    # SURVEY.md App. A.4 (derived independently by the surveyor, 6 printed digits)
    sim = oe.OracleSim()
    sim.set_joint_angles([0, -1.5708, 0, -1.5708, 0, 0])
    np.testing.assert_allclose(sim.get_link_position(7), [-0.000673, -0.232972, 1.080159], atol=1e-6)
    np.testing.assert_allclose(sim.get_link_orientation(7), [-1.571051, 0.000007, 3.138260], atol=1e-6)
    sim.set_joint_angles(np.zeros(6))
    np.testing.assert_allclose(sim.get_link_position(7), [-0.817267, -0.234444, 0.062675], atol=1e-6)
    np.testing.assert_allclose(sim.get_link_orientation(7), [1.571389, -0.000010, 0.001307], atol=1e-6)


def test_euler_quaternion_roundtrip_and_scipy_agreement():
    from scipy.spatial.transform import Rotation as R
    sim = oe.OracleSim()
    rng = np.random.default_rng(1)
    for e in rng.uniform(-np.pi, np.pi, (200, 3)) * [1, 0.49, 1]:
        q = sim.euler_to_quaternion(e)
        # pybullet (roll,pitch,yaw) == scipy extrinsic 'xyz'
        np.testing.assert_allclose(np.abs(np.dot(q, R.from_euler("xyz", e).as_quat())), 1.0, atol=1e-14)
        sim.add_body("b", "box"); sim.set_base_pose("b", np.zeros(3), q)
        np.testing.assert_allclose(sim.get_base_rotation("b"), e, atol=1e-9)
    # |pitch| > 90 deg reads back on the other Euler branch but is the same rotation (quirk Q3)
    e = np.deg2rad([100.0, 120.0, 0.0])
    q = sim.euler_to_quaternion(e)
    sim.set_base_pose("b", np.zeros(3), q)
    back = sim.get_base_rotation("b")
    assert not np.allclose(back, e)
    np.testing.assert_allclose(np.abs(np.dot(sim.euler_to_quaternion(back), q)), 1.0, atol=1e-12)


def test_quaternion_difference_axis_angle():
    from scipy.spatial.transform import Rotation as R
    sim = oe.OracleSim()
    rng = np.random.default_rng(2)
    for _ in range(100):
        a, b = R.random(random_state=rng.integers(1 << 30)), R.random(random_state=rng.integers(1 << 30))
        d = sim.get_quaternion_difference(a.as_quat(), b.as_quat())
        axis, angle = sim.get_axis_angle(d)
        assert 0 <= angle <= np.pi + 1e-12           # nearest() keeps the short way round
        rel = R.from_rotvec(np.array(axis) * angle)
        np.testing.assert_allclose((rel * a).as_matrix(), b.as_matrix(), atol=1e-12)


def _sep_max(supA, supB, n0):
    """max over unit n of  min_a n.a - max_b n.b   (= distance when positive)."""
    def f(x):
        n = x / np.linalg.norm(x)
        return -(-supA(-n) - supB(n))
    best = None
    starts = [n0] + [np.eye(3)[i] * sg + 1e-3 for i in range(3) for sg in (1, -1)]
    starts.sort(key=f)
    for s in starts[:3]:
        r = minimize(f, s, method="Nelder-Mead", options={"xatol": 1e-11, "fatol": 1e-14, "maxiter": 20000, "maxfev": 20000})
        if best is None or r.fun < best.fun:
            best = r
    return -best.fun


def _hull_support(link, pos, rot):
    v = MODEL["hull_vertices"][MODEL["hull_offsets"][link]:MODEL["hull_offsets"][link + 1]] @ rot[link].T + pos[link]
    return (lambda n: np.max(v @ n)), v.mean(0)


def _box_support(c, he):
    c, he = np.array(c), np.array(he)
    return (lambda n: c @ n + np.abs(n) @ he), c


def _cyl_support(c, R_, r, h):
    def s(n):
        l = R_.T @ n
        return c @ n + r * np.hypot(l[0], l[1]) + h * abs(l[2])
    return s, np.array(c)


@pytest.mark.parametrize("seed", range(6))
def test_gjk_distance_equals_dual_separation(seed):
    rng = np.random.default_rng(100 + seed)
    sim = oe.OracleSim()
    sim.add_body("plane", "box"); sim.add_body("table", "box"); sim.add_body("track", "box")
    sim.add_body("target", "box"); sim.add_body("obstacle", "cylinder")
    q = np.array([0, -1.5708, 0, -1.5708, 0, 0]) + rng.uniform(-1.2, 1.2, 6)
    sim.set_joint_angles(q)
    opos = rng.uniform([0.3, -0.5, 0.2], [0.9, 0.5, 0.7]); oeul = rng.uniform(-2.5, 2.5, 3)
    sim.set_base_pose("obstacle", opos, oeul)
    pos, rot = oracle_fk(q)
    sc = sim._scene()
    from scipy.spatial.transform import Rotation as R
    Ro = R.from_quat(sim.euler_to_quaternion(oeul)).as_matrix()
    m = 0.001
    others = {0: _box_support([0.5, 0, -0.58], [0.55 - m, 0.9 - m, 0.46 - m]),
              1: _box_support([0, 0, -0.06], [0.1 - m, 0.55 - m, 0.06 - m]),
              2: _cyl_support(opos, Ro, 0.05 - m, 0.2 - m)}
    checked = 0
    for link in range(1, 7):
        supL, cL = _hull_support(link, pos, rot)
        pairs = [(o, others[o]) for o in (0, 1, 2)] + [(3 + l2, _hull_support(l2, pos, rot)) for l2 in range(link + 2, 7)]
        for oid, (supO, cO) in pairs:
            deep, iters = ctypes.c_int(), ctypes.c_int()
            d = oe.lib().orc_pair_distance(ctypes.byref(sc), link, oid, ctypes.byref(deep), ctypes.byref(iters))
            if deep.value:
                continue
            marg = 0.002
            core = d + marg
            want = _sep_max(supL, supO, cL - cO)
            assert abs(core - want) < 2e-8, (link, oid, core, want, iters.value)
            checked += 1
    assert checked >= 10
