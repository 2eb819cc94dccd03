"""CPU tier of the closed-loop check (tests/closed_loop.py): the product's per-env code compiled for the host
(tests/hostcheck, capsule geometry), the shipped policies, natural-reset scenarios, 400 episodes per task -- a coarse
version of tests/test_gpu_closed_loop.py (binomial standard error 1-2 points) that runs without a GPU."""
import pytest

from tests.closed_loop import POLICY_LINK_DIST, run_host

ALL = ["UR5OriReach-v1", "UR5ObsReach-v1", "UR5StaReach-v1", "UR5DynReach-v1"]


@pytest.mark.parametrize("env_id", ALL)
def test_shipped_policy_on_host_instantiation(env_id):
    r = run_host(env_id, geom=1, n=400, link_dist_mode=1 if POLICY_LINK_DIST[env_id] == "workbench" else 0)
    assert abs(r["success_rate_pct"] - r["published"]["published_success_rate_pct"]) <= 6.0, r


def test_obstacle_only_link_dist_breaks_the_2023_policies():
    r = run_host("UR5ObsReach-v1", geom=1, n=300, link_dist_mode=0)
    assert r["collision_rate_pct"] >= 20.0, r
    z = run_host("UR5ObsReach-v1", geom=1, n=300, link_dist_mode=0, ld_ablation="zero")
    assert z["collision_rate_pct"] <= 12.0 and z["success_rate_pct"] >= r["success_rate_pct"] + 15.0, (r, z)
