"""Closed-loop statistical pin against numbers real PyBullet produced: the reference's
shipped SAC policies, run through the reference's model_test.py protocol on this simulator (hull geometry = the
reference's geometry), against the success rates the reference publishes for them (Trained_Models/*/best.txt).
See tests/closed_loop.py for the protocol and for the link-distance definition each policy is run with.

All four tasks are asserted within 3 points of the published success rate.  UR5ObsReach / UR5StaReach additionally
pin the time-out and collision shares of the published per-episode lines, and the test keeps the round-1 failure
visible: with the link-distance definition of the SHIPPED code (obstacle only) the two 2023 policies must still fail the
way they did (the table dive / the stalls), and setting the policies' link_dist inputs to zero must repair most of it --
if either stops being true, the diagnosis in DESIGN.md section 2 is wrong and has to be redone."""
import pytest

pytestmark = pytest.mark.gpu
ALL = ["UR5OriReach-v1", "UR5ObsReach-v1", "UR5StaReach-v1", "UR5DynReach-v1"]


@pytest.mark.parametrize("env_id", ALL)
def test_shipped_policy_success_rate_matches_published(env_id):
    from tests.closed_loop import run
    r = run(env_id, "hull")
    pub = r["published"]
    assert abs(r["success_rate_pct"] - pub["published_success_rate_pct"]) <= 3.0, r
    assert abs(r["timeout_pct"] - pub["published_timeout_pct"]) <= 2.0, r
    assert abs(r["collision_rate_pct"] - pub["published_collision_pct"]) <= 2.0, r
    assert abs(r["success_mean_steps"] - pub["published_success_mean_steps"]) <= 0.5, r
    if env_id != "UR5StaReach-v1":
        # UR5StaReach's published reward sums are 28 lower per successful episode than any variant here produces at
        # the same step counts: its 2023 reward weights are not the shipped ones (not recoverable from the artefacts)
        assert abs(r["mean_reward"] - pub["published_mean_reward"]) <= 0.1 * abs(pub["published_mean_reward"]), r
        assert abs(r["mean_steps"] - pub["published_mean_steps"]) <= 1.0, r
    if env_id in ("UR5OriReach-v1", "UR5DynReach-v1"):
        # the scenario list is a grid of goal positions in a fixed order: the spatial profile of steps and failures
        from tests.closed_loop import compare_profiles
        compare_profiles(r)


@pytest.mark.parametrize("env_id", ALL)
def test_capsule_geometry_same_statistics(env_id):
    """the throughput path's calibrated capsules give the same closed-loop statistics as the hulls"""
    from tests.closed_loop import run
    r = run(env_id, "capsule")
    assert abs(r["success_rate_pct"] - r["published"]["published_success_rate_pct"]) <= 3.0, r


@pytest.mark.parametrize("env_id", ["UR5ObsReach-v1", "UR5StaReach-v1"])
def test_diagnosis_stays_reproducible(env_id):
    from tests.closed_loop import run
    shipped = run(env_id, "hull", link_dist="obstacle")
    zeroed = run(env_id, "hull", link_dist="obstacle", ld_ablation="zero")
    pub = shipped["published"]["published_success_rate_pct"]
    assert shipped["success_rate_pct"] <= pub - 20.0, shipped          # round 1: 65.2 / 63.7
    assert zeroed["success_rate_pct"] >= shipped["success_rate_pct"] + 15.0, (shipped, zeroed)
    if env_id == "UR5ObsReach-v1":
        assert shipped["collision_rate_pct"] >= 20.0 and zeroed["collision_rate_pct"] <= 10.0, (shipped, zeroed)
    else:
        assert shipped["timeout_pct"] >= 20.0 and zeroed["timeout_pct"] <= 10.0, (shipped, zeroed)
