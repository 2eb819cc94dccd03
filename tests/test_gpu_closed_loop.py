"""Closed-loop statistical pin against numbers real PyBullet produced: the reference's
shipped SAC policies, run through the reference's model_test.py protocol on this simulator, against the success rates
the reference publishes for them (Trained_Models/*/best.txt).  See tests/closed_loop.py for the protocol.

Status (profiles/closed_loop_r01.json, hull geometry = the reference's geometry):
  UR5OriReach-v1  96.9 % here vs 97.3 % published;  UR5DynReach-v1  96.9 % vs 96.2 %   -> asserted within 3 points
  UR5ObsReach-v1  65 % vs 95.9 %;  UR5StaReach-v1  64 % vs 89.4 %                          -> KNOWN GAP, see DESIGN.md
The two tasks with a static obstacle above the goal region do not reproduce: in Obs the policy rests wrist_1 / the
forearm on the table for low goals (ignoring robot-table pairs gives 90 %), in Sta it converges to fixed points
1-3 cm outside the success thresholds.  Bullet's contact response (live robot-table / robot-target contacts) is
the part of the reference's physics that the hold-pose model does not restate; the gap is recorded,
bounded from below here so that it cannot silently grow, and not claimed as parity."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env_id", ["UR5OriReach-v1", "UR5DynReach-v1"])
def test_shipped_policy_success_rate_matches_published(env_id):
    from tests.closed_loop import run
    r = run(env_id, "hull")
    pub = r["published"]
    assert abs(r["success_rate_pct"] - pub["published_success_rate_pct"]) <= 3.0, r
    assert abs(r["mean_reward"] - pub["published_mean_reward"]) <= 0.1 * abs(pub["published_mean_reward"]), r
    assert abs(r["mean_steps"] - pub["published_mean_steps"]) <= 1.0, r


@pytest.mark.parametrize("env_id,floor", [("UR5ObsReach-v1", 55.0), ("UR5StaReach-v1", 55.0)])
def test_known_gap_does_not_grow(env_id, floor):
    from tests.closed_loop import run
    r = run(env_id, "hull")
    assert r["success_rate_pct"] >= floor, r
    assert r["success_rate_pct"] <= r["published"]["published_success_rate_pct"] + 3.0, r
