"""Tier T0: the oracle's restatement of UR_gym/utils.py against outputs of the reference module itself
(tests/golden/utils_golden.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np

from oracle import oracle_env as oe


def test_distance_matches_reference(golden):
    a, b = golden["a"], golden["b"]
    np.testing.assert_allclose(oe.distance(a, b), golden["dist_batched"], rtol=0, atol=1e-15)
    rows = np.array([oe.distance(a[i], b[i])[0] for i in range(len(a))])
    np.testing.assert_allclose(rows, golden["dist_rows"], rtol=0, atol=1e-15)
    d3 = np.array([oe.distance(a[i, :3], b[i, :3])[0] for i in range(len(a))])
    np.testing.assert_allclose(d3, golden["dist3"], rtol=0, atol=1e-15)
    # float32 achieved goal against float64 goal, as RobotTaskEnv.step feeds them (core.py:256,313)
    a32 = a.astype(np.float32)
    f32 = np.array([oe.distance(a32[i], b[i])[0] for i in range(len(a))])
    np.testing.assert_allclose(f32, golden["dist_f32"], rtol=0, atol=1e-15)


def test_angular_distance_matches_reference(golden):
    a, b = golden["a"], golden["b"]
    # 2*arccos(|dot|) amplifies round-off near dot = 1 (d angle / d dot = 2/sin(angle/2)): the reference's own
    # value for identical rotations is only accurate to ~sqrt(eps).  Compare on the dot product scale.
    for got, want in ((oe.angular_distance(a, b), golden["ang_batched"]),
                      (np.array([oe.angular_distance(a[i], b[i])[0] for i in range(len(a))]), golden["ang_rows"]),
                      (np.array([oe.angular_distance(a.astype(np.float32)[i], b[i])[0] for i in range(len(a))]),
                       golden["ang_f32"])):
        np.testing.assert_allclose(np.cos(got / 2), np.cos(want / 2), rtol=0, atol=4e-16)
        big = want > 1e-3
        np.testing.assert_allclose(got[big], want[big], rtol=0, atol=1e-12)
        assert np.all(np.abs(got[~big] - want[~big]) < 1e-7)


def test_sample_euler_constrained_matches_reference(golden):
    got = np.array([oe.euler_constrained_from_uniform(*u) for u in golden["cons_u"]])
    np.testing.assert_allclose(got, golden["cons"], rtol=0, atol=1e-15)
    assert np.all(got[:, 0] <= np.deg2rad(-90)) and np.all(got[:, 0] > np.deg2rad(-180) - 1e-12)
    assert np.all(got[:, 1] == 0) and np.all(got[:, 2] <= 0) and np.all(got[:, 2] > -np.pi - 1e-12)


def test_sample_euler_obstacle_matches_reference(golden):
    got = np.array([oe.euler_obstacle_from_uniform(*u) for u in golden["obst_u"]])
    np.testing.assert_allclose(got, golden["obst"], rtol=0, atol=1e-15)
    assert np.all(got[:, 2] == 0)
    assert set(np.sign(got[:, 0])) == {-1.0, 1.0}
