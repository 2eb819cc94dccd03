"""GPU tier: the CUDA library, called through its C ABI (ctypes) by the Python host layer, against the oracle.

Same harness as the CPU tier (tests/_parity.py): identical actions, identical counter-based reset stream, every
observation / reward / flag / terminal observation / post-reset observation of every step compared at the
north-star tolerances.  Then size-independent properties at BASELINE.json's full size (1 Mi envs per GPU)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import oracle_env as oe
from tests._parity import obs_close, run_parity, sta_moving_scenarios

pytestmark = pytest.mark.gpu
TASKS = ["UR5OriReach-v1", "UR5ObsReach-v1", "UR5StaReach-v1", "UR5DynReach-v1"]
# hull geometry: FP32 link poses differ from the oracle's FP64 poses by up to ~1.5e-6 m, and so do the GJK distances
# built on them (pairs whose FP32 iteration does not converge are redone in FP64: urgym_device.cuh gjk_distance_refine).
# The budget: link distances 5e-6 m; rewards additionally 100 (Obs weight) x 5 links x 2 distances x 2e-6 m.
HULL_LD_TOL, HULL_REW_ATOL = 5e-6, 2e-3


@pytest.fixture(scope="module")
def ug():
    import urgym_b200
    return urgym_b200


@pytest.mark.parametrize("env_id", TASKS)
def test_capsule_rollout_parity(env_id, ug):
    from tests._gpu import GpuSim
    n, steps = 200, 110        # 200 = one full 128-row tile + a ragged one; 110 steps cross the TimeLimit
    st = run_parity(GpuSim(env_id, oe.GEOM_CAPSULE, n, seed=7), env_id, oe.GEOM_CAPSULE, n, steps, seed=7)
    assert st["steps"] > 0.9 * n * steps and st["resets"] > 0, st
    print(env_id, st)


@pytest.mark.parametrize("env_id", TASKS)
def test_capsule_small_actions_reach_timelimit(env_id, ug):
    """small actions keep the arm away from collisions so episodes run into the 100-step TimeLimit"""
    from tests._gpu import GpuSim
    n, steps = 64, 205
    st = run_parity(GpuSim(env_id, oe.GEOM_CAPSULE, n, seed=11, offset=5000), env_id, oe.GEOM_CAPSULE, n, steps, seed=11,
                    offset=5000, action_scale=0.05)
    assert st["truncations"] > 0, st


@pytest.mark.parametrize("env_id", TASKS)
def test_hull_rollout_parity(env_id, ug):
    from tests._gpu import GpuSim
    n, steps = 40, 40
    st = run_parity(GpuSim(env_id, oe.GEOM_HULL, n, seed=3, offset=77), env_id, oe.GEOM_HULL, n, steps, seed=3, offset=77,
                    ld_tol=HULL_LD_TOL, rew_atol=HULL_REW_ATOL)
    assert st["steps"] > 0.85 * n * steps, st


@pytest.mark.parametrize("env_id", TASKS)
def test_gpu_equals_host_instantiation_bitwise_flags(env_id, ug):
    """the kernel and tests/hostcheck run the same functions; booleans must agree except next to a threshold, floats
    to FP32 round-off (FMA contraction differs between nvcc and g++)"""
    from tests._gpu import GpuSim
    from tests._hostcheck import HostCheckSim
    n, steps = 1000, 30
    g, h = GpuSim(env_id, 1, n, seed=5), HostCheckSim(env_id, 1, n, seed=5)
    og, oh = g.reset(), h.reset()
    np.testing.assert_allclose(og[:, :3], oh[:, :3], atol=2e-6)          # Euler columns wrap at +-pi: see the oracle tests
    np.testing.assert_allclose(og[:, 6:15], oh[:, 6:15], atol=2e-6)
    rng = np.random.default_rng(0)
    alive = np.ones(n, bool)
    for t in range(steps):
        a = rng.uniform(-1.1, 1.1, (n, 6)).astype(np.float32)
        rg, rh = g.step(a), h.step(a)
        same = (rg["terminated"] == rh["terminated"]) & (rg["truncated"] == rh["truncated"]) & (rg["is_success"] == rh["is_success"])
        alive &= same
        # positions, joints, goal (Euler columns wrap at +-pi and are covered by the oracle tests); the GPU uses the
        # SFU for sin / cos (abs error 3.6e-7), the host instantiation libm.  Finished envs: terminal rows.
        done = rh["terminated"] | rh["truncated"]
        for sel, key in ((alive & done, "terminal_obs"), (alive & ~done, "obs")):
            np.testing.assert_allclose(rg[key][sel][:, :3], rh[key][sel][:, :3], atol=5e-6)
            np.testing.assert_allclose(rg[key][sel][:, 6:15], rh[key][sel][:, 6:15], atol=5e-6)
        np.testing.assert_allclose(rg["reward"][alive], rh["reward"][alive], rtol=1e-5, atol=2e-3)
    assert alive.mean() > 0.99


def test_sharding_invariance(ug):
    """the same global env index produces the same episodes whatever the shard layout"""
    env_id, n = "UR5DynReach-v1", 1024
    whole = ug.UR5VecEnv(env_id, n, seed=42)
    parts = [ug.UR5VecEnv(env_id, hi - lo, seed=42, env_index_offset=lo)
             for lo, hi in (ug.shard_range(n, r, 3) for r in range(3))]
    o = whole.reset()["observation"].clone()
    op = torch.cat([p.reset()["observation"] for p in parts])
    assert torch.equal(o, op)
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(40):
        a = torch.rand((n, 6), device="cuda", generator=g) * 2 - 1
        ow, rw, tw, cw, iw = whole.step(a)
        outs, lo = [], 0
        for p in parts:
            outs.append(p.step(a[lo:lo + p.num_envs].contiguous())); lo += p.num_envs
        assert torch.equal(ow["observation"], torch.cat([x[0]["observation"] for x in outs]))
        assert torch.equal(rw, torch.cat([x[1] for x in outs]))
        assert torch.equal(tw, torch.cat([x[2] for x in outs]))
        assert torch.equal(cw, torch.cat([x[3] for x in outs]))
    sw = whole.stats()
    sp = [p.stats() for p in parts]
    for k in sw:
        assert sw[k] == sum(s[k] for s in sp), k      # fixed-point return sum: exact, order-independent
    assert sw["episodes"] > 0 and sw["env_steps"] == n * 40


@pytest.mark.parametrize("chains", [2, 3, 8])
def test_chained_graph_equals_plain_steps(ug, chains):
    """a CUDA graph that advances env sub-ranges as independent chains == the same steps issued one by one"""
    env_id, n, k = "UR5DynReach-v1", 50_000, 6
    a_env, b_env = ug.UR5VecEnv(env_id, n, seed=13), ug.UR5VecEnv(env_id, n, seed=13)
    a_env.reset(); b_env.reset()
    g = torch.Generator(device="cuda").manual_seed(2)
    ring = [torch.rand((n, 6), device="cuda", generator=g) * 2 - 1 for _ in range(k)]
    graph = b_env.capture_steps(ring, chains=chains)
    for rep in range(4):
        for a in ring:
            a_env.step(a)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(a_env.obs, b_env.obs) and torch.equal(a_env.reward, b_env.reward)
        assert torch.equal(a_env.terminated, b_env.terminated) and torch.equal(a_env.truncated, b_env.truncated)
    for name in ("q", "goal", "obstacle", "obstacle_end", "link_dist", "elapsed", "ep_return"):
        assert torch.equal(a_env.get_state(name), b_env.get_state(name)), name
    sa, sb = a_env.stats(), b_env.stats()
    assert sa == sb and sa["episodes"] > 0
    # a whole-batch step after the chains keeps the event numbering consistent
    a_env.step(ring[0]); b_env.step(ring[0])
    assert torch.equal(a_env.obs, b_env.obs)


def test_fused_autoreset_equals_explicit_reset(ug):
    """step with auto-reset == step without it followed by reset(mask = done) at the same reset event"""
    env_id, n = "UR5StaReach-v1", 4096
    a_env = ug.UR5VecEnv(env_id, n, seed=9)
    b_env = ug.UR5VecEnv(env_id, n, seed=9, auto_reset=False)
    a_env.reset(); b_env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in range(25):
        act = torch.rand((n, 6), device="cuda", generator=g) * 2.4 - 1.2
        oa, ra, ta, ca, ia = a_env.step(act)
        ob, rb, tb, cb, ib = b_env.step(act)
        assert torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(ca, cb)
        done = (tb | cb)
        assert torch.equal(ia["terminal_observation"][done.bool()], ob["observation"][done.bool()])
        ev = ctypes.c_uint32()
        b_env.L.urgym_get_event(b_env.h, ctypes.byref(ev))
        b_env.L.urgym_set_event(b_env.h, ev.value - 1)
        ob2 = b_env.reset(mask=done)
        assert torch.equal(oa["observation"], ob2["observation"])
    for name in ("q", "goal", "obstacle", "link_dist", "elapsed", "ep_return"):
        assert torch.equal(a_env.get_state(name), b_env.get_state(name)), name


def test_injected_episode_constants_reach_the_step_kernel(ug):
    """urgym_set_state of goal / obstacle fields rebuilds the hot planes (episode cache) the step kernel reads:
    two handles that differ only in HOW they got their episode constants (own reset vs injection) step identically"""
    for env_id in ("UR5OriReach-v1", "UR5ObsReach-v1", "UR5StaReach-v1", "UR5DynReach-v1"):
        n = 777
        a_env = ug.UR5VecEnv(env_id, n, seed=21, auto_reset=False)
        b_env = ug.UR5VecEnv(env_id, n, seed=99, auto_reset=False)
        a_env.reset(); b_env.reset()
        names = ["goal"] + (["obstacle"] if env_id != "UR5OriReach-v1" else []) + (["obstacle_end"] if env_id == "UR5DynReach-v1" else [])
        for name in names:
            b_env.set_state(name, a_env.get_state(name))
        b_env.refresh()
        if env_id != "UR5OriReach-v1":
            # (reset takes the neutral pose's capsules from the host's double-precision FK, refresh runs the FP32 chain)
            assert torch.allclose(a_env.get_state("link_dist"), b_env.get_state("link_dist"), atol=2e-6)
            b_env.set_state("link_dist", a_env.get_state("link_dist"))
        g = torch.Generator(device="cuda").manual_seed(5)
        for t in range(6):
            act = torch.rand((n, 6), device="cuda", generator=g) * 2 - 1
            oa, ra, ta, ca, _ = a_env.step(act)
            ob, rb, tb, cb, _ = b_env.step(act)
            # the cache words come from two different kernels (reset vs derive): the same formulas, but the compiler
            # may contract them into FMAs differently, so equality is to round-off, not bitwise
            assert torch.allclose(oa["observation"], ob["observation"], atol=3e-6, rtol=0), (env_id, t)
            assert torch.allclose(ra, rb, atol=2e-3, rtol=1e-5) and torch.equal(ta, tb) and torch.equal(ca, cb)


def test_kernel_timing_entry_points(ug):
    """urgym_profile_enable / urgym_profile_read: CUDA events around the step and auto-reset kernels"""
    env = ug.UR5VecEnv("UR5DynReach-v1", 65536, seed=1)
    env.reset()
    act = torch.zeros((65536, 6), device="cuda")
    assert env.L.urgym_profile_enable(env.h, 1) == 0
    for _ in range(5):
        env.step(act)
    a, b, k = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    assert env.L.urgym_profile_read(env.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(k)) == 0
    assert k.value == 5 and 0.0 < a.value < 5.0 and 0.0 < b.value < 5.0
    assert env.L.urgym_profile_enable(env.h, 0) == 0
    env.step(act)
    assert env.L.urgym_profile_read(env.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(k)) == 0 and k.value == 0


def test_sta_4mi_envs_sharded_properties(ug):
    """BASELINE config C4's size: UR5StaReach-v1, 4 Mi envs, here as two shards of 2 Mi on one GPU.  Size-independent
    properties: sampling boxes, target-obstacle rejection rule, determinism of a shard against the same shard re-run,
    statistics additive and equal to a torch-side recount."""
    env_id, n_shard = "UR5StaReach-v1", 1 << 21
    shards = [ug.UR5VecEnv(env_id, n_shard, seed=77, env_index_offset=k * n_shard) for k in range(2)]
    again = ug.UR5VecEnv(env_id, n_shard, seed=77, env_index_offset=n_shard)
    for e in shards + [again]:
        e.reset()
    lo, hi = torch.tensor([0.3, -0.5, 0.0], device="cuda"), torch.tensor([0.75, 0.5, 0.2], device="cuda")
    olo, ohi = torch.tensor([0.5, -0.5, 0.25], device="cuda"), torch.tensor([1.0, 0.5, 0.55], device="cuda")
    for e in shards:
        goal, obst = e.get_state("goal"), e.get_state("obstacle")
        assert ((goal[:, :3] >= lo) & (goal[:, :3] <= hi)).all()                    # reach.py:385-386
        assert ((obst[:, :3] >= olo) & (obst[:, :3] <= ohi)).all()                  # reach.py:387-388
        # rejection rule (reach.py:473): the capsule stand-in keeps target and obstacle axis farther apart than
        # 0.1 + the two radii minus the obstacle's half length along its axis
        assert ((goal[:, :3] - obst[:, :3]).norm(dim=1) > 0.1).all()
    assert not torch.equal(shards[0].get_state("goal"), shards[1].get_state("goal"))
    g = torch.Generator(device="cuda").manual_seed(11)
    done_count, succ_count = [0, 0], [0, 0]
    for t in range(12):
        act = torch.rand((n_shard, 6), device="cuda", generator=g) * 2 - 1
        for k, e in enumerate(shards):
            obs, rew, term, trunc, info = e.step(act)
            assert torch.isfinite(rew).all()
            done_count[k] += int((term | trunc).sum()); succ_count[k] += int(info["is_success"].sum())
        o2, r2, t2, c2, _ = again.step(act)
        assert torch.equal(o2["observation"], shards[1].obs) and torch.equal(r2, shards[1].reward) and torch.equal(t2, shards[1].terminated)
    st = [e.stats() for e in shards]
    for k in range(2):
        assert st[k]["episodes"] == done_count[k] and st[k]["successes"] == succ_count[k]
        assert st[k]["env_steps"] == 12 * n_shard
    assert sum(s["episodes"] for s in st) > 0


def test_full_size_properties(ug):
    """BASELINE config: UR5DynReach-v1, 1 Mi envs on one GPU.  Size-independent properties."""
    env_id, n, steps = "UR5DynReach-v1", 1 << 20, 120
    env = ug.UR5VecEnv(env_id, n, seed=2026)
    first = env.reset()["observation"].clone()
    assert torch.isfinite(first).all()
    # goals / obstacles inside the sampling boxes (reach.py:584-587), start-end at least 1 m apart (reach.py:674-675)
    goal, start, end = env.get_state("goal"), env.get_state("obstacle"), env.get_state("obstacle_end")
    lo, hi = torch.tensor([0.4, -0.5, 0.0], device="cuda"), torch.tensor([0.75, 0.5, 0.2], device="cuda")
    assert ((goal[:, :3] >= lo) & (goal[:, :3] <= hi)).all()
    olo, ohi = torch.tensor([0.5, -0.8, 0.25], device="cuda"), torch.tensor([1.2, 0.8, 0.75], device="cuda")
    assert ((start[:, :3] >= olo) & (start[:, :3] <= ohi) & (end[:, :3] >= olo) & (end[:, :3] <= ohi)).all()
    assert ((end[:, :3] - start[:, :3]).norm(dim=1) >= 1.0 - 1e-6).all()
    g = torch.Generator(device="cuda").manual_seed(3)
    ep_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    n_done = n_succ = n_trunc = 0
    ret_sum = 0.0
    ep_ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    small = torch.arange(n, device="cuda") % 4 == 0        # a quarter of the envs take tiny actions -> hit the TimeLimit
    for t in range(steps):
        a = torch.rand((n, 6), device="cuda", generator=g) * 2 - 1
        a[small] *= 0.03
        obs, rew, term, trunc, info = env.step(a)
        assert torch.isfinite(obs["observation"]).all() and torch.isfinite(rew).all()
        ep_len += 1
        ep_ret += rew.double()
        done = (term | trunc).bool()
        assert not (info["is_success"].bool() & ~term.bool()).any()            # success implies terminated (core.py:313-315)
        assert torch.equal(trunc.bool(), ep_len >= 100)                          # TimeLimit
        assert (env.get_state("elapsed")[~done] == ep_len[~done]).all()
        # a finished env restarts at the neutral pose with a fresh goal
        q = obs["observation"][:, 6:12]
        neutral = torch.tensor([0.0, -1.5708, 0.0, -1.5708, 0.0, 0.0], device="cuda")
        assert (q[done] == neutral).all()
        n_done += int(done.sum()); n_succ += int(info["is_success"].sum()); n_trunc += int((trunc.bool() & ~term.bool()).sum())
        ret_sum += float(ep_ret[done].sum())
        ep_len[done] = 0; ep_ret[done] = 0
    st = env.stats()
    assert st["env_steps"] == n * steps and st["episodes"] == n_done and st["successes"] == n_succ and st["truncations"] == n_trunc
    assert n_trunc > 0
    assert abs(st["return_sum"] - ret_sum) <= 1e-4 * abs(ret_sum) + 1.0
    # determinism: the same seed and actions reproduce the run bit for bit
    env2 = ug.UR5VecEnv(env_id, n, seed=2026)
    assert torch.equal(env2.reset()["observation"], first)


def test_state_roundtrip_and_checkpoint(ug):
    env_id, n = "UR5DynReach-v1", 3000
    env = ug.UR5VecEnv(env_id, n, seed=1)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = [torch.rand((n, 6), device="cuda", generator=g) * 2 - 1 for _ in range(30)]
    for a in acts[:10]:
        env.step(a)
    ck = env.state_dict()
    ref = [tuple(x.clone() if torch.is_tensor(x) else x for x in (env.step(a)[0]["observation"], env.reward)) for a in acts[10:]]
    env2 = ug.UR5VecEnv(env_id, n, seed=1)
    env2.load_state_dict(ck)
    for a, (o, r) in zip(acts[10:], ref):
        o2 = env2.step(a)[0]["observation"]
        assert torch.equal(o2, o) and torch.equal(env2.reward, r)
    for name in ("q", "goal", "obstacle", "obstacle_end", "link_dist", "velocity", "ep_return"):
        v = torch.randn_like(env.get_state(name))
        env.set_state(name, v)
        assert torch.equal(env.get_state(name), v), name
    with pytest.raises(ug.UrgymError):
        ug.UR5VecEnv("UR5OriReach-v1", 8).get_state("obstacle")


@pytest.mark.parametrize("n", [5000, 600_000])       # 600 k: the chunked three-stream pipeline (8 chunks)
def test_host_buffer_entry_point_matches_device_path(ug, n):
    env_id = "UR5ObsReach-v1"
    d_env, h_env = ug.UR5VecEnv(env_id, n, seed=4), ug.UR5VecEnv(env_id, n, seed=4)
    d_env.reset(); h_env.reset()
    buf = h_env.alloc_host_buffers()
    rng = np.random.default_rng(0)
    for t in range(12):
        a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        buf["actions"].copy_(torch.from_numpy(a))
        h_env.step_host(buf["actions"], buf)
        o, r, te, tr, info = d_env.step(torch.from_numpy(a).cuda())
        assert torch.equal(buf["obs"], o["observation"].cpu()) and torch.equal(buf["reward"], r.cpu())
        assert torch.equal(buf["terminated"], te.cpu()) and torch.equal(buf["truncated"], tr.cpu())
        done = (te | tr).bool().cpu()
        assert torch.equal(buf["terminal_obs"][done], info["terminal_observation"].cpu()[done])


@pytest.mark.parametrize("env_id", TASKS)
def test_gym_style_single_env_matches_oracle(env_id, ug):
    """demo.py's loop (reset on done) through make()/reset()/step() against one oracle env"""
    env = ug.make(env_id, render=False, seed=21)
    orc = oe.make(env_id, geom=oe.GEOM_CAPSULE, stream=oe.PhiloxStream(21), env_index=0, first_event=1)
    assert env.observation_space["observation"].shape == (ug.UR5VecEnv._FIELDS and env.vec.obs_dim,)
    assert env.action_space.shape == (6,)
    event = 2                                    # constructor reset = event 1, explicit reset below = event 2
    obs, info = env.reset()
    o2, i2 = orc.reset(event=event)
    np.testing.assert_allclose(obs["observation"], o2["observation"], atol=1e-5)
    np.testing.assert_allclose(obs["desired_goal"], o2["desired_goal"], atol=1e-6)
    rng = np.random.default_rng(0)
    for t in range(150):
        a = rng.uniform(-1, 1, 6).astype(np.float32)
        obs, r, term, trunc, info = env.step(a); event += 1
        o2, r2, term2, trunc2, info2 = orc.step(a)
        assert (term, trunc, info["is_success"]) == (term2, trunc2, bool(info2["is_success"]))
        np.testing.assert_allclose(obs["achieved_goal"][:3], o2["achieved_goal"][:3], atol=1e-5)
        assert abs(r - r2) <= 1e-5 * max(1.0, abs(r2))
        # the reference's own API pieces evaluated on the returned arrays
        assert bool(np.all(env.task.is_success(obs["achieved_goal"], obs["desired_goal"]))) == bool(np.all(orc.task.is_success(o2["achieved_goal"], orc.task.get_goal())))
        rr = float(np.asarray(env.compute_reward(obs["achieved_goal"], obs["desired_goal"], info)).reshape(-1)[0])
        assert abs(rr - r2) <= 2e-4 * max(1.0, abs(r2)), (rr, r2)
        if term or trunc:
            obs, info = env.reset(); event += 1
            o2, i2 = orc.reset(event=event)
            lin, ang = obs_close(env_id, obs["observation"], o2["observation"])
            assert lin <= 1e-5 and ang <= 1e-5
    env.close()


def test_injection_protocol_like_model_test(ug):
    """model_test.py:34-38: reset, inject goal + obstacle, read the observation back, step"""
    env = ug.make("UR5DynReach-v1", seed=0)
    orc = oe.make("UR5DynReach-v1", geom=oe.GEOM_CAPSULE, stream=oe.PhiloxStream(0), env_index=0, first_event=1)
    env.reset(); orc.reset(event=2)
    data = np.array([0.5, 0.1, 0.1, -2.0, 0.0, -1.0, 0.7, -0.6, 0.4, 1.0, 0.8, 0.0, 0.9, 0.7, 0.6, -2.0, -1.2, 0.0])
    env.task.set_goal_and_obstacle(data); orc.task.set_goal_and_obstacle(data)
    np.testing.assert_allclose(env.task.get_obs(), orc.task.get_obs(), atol=1e-5)
    np.testing.assert_allclose(env.robot.get_obs(), orc.robot.get_obs(), atol=1e-5)
    np.testing.assert_allclose(env.task.link_dist, orc.task.link_dist, atol=1e-5)
    rng = np.random.default_rng(1)
    for t in range(40):
        a = rng.uniform(-0.3, 0.3, 6).astype(np.float32)
        o, r, term, trunc, info = env.step(a)
        o2, r2, term2, trunc2, info2 = orc.step(a)
        lin, ang = obs_close("UR5DynReach-v1", o["observation"], o2["observation"])
        assert lin <= 1e-5 and ang <= 1e-5
        assert abs(r - r2) <= 1e-5 * max(1.0, abs(r2)) and term == term2
        if term:
            break


def test_c_abi_error_behaviour(ug):
    L = ug._native.lib()
    h = ctypes.c_void_p()
    assert L.urgym_create(ctypes.byref(h), 9, 1, 4, 0, 0, 0) == -1 and b"task" in L.urgym_last_error(None)
    assert L.urgym_create(ctypes.byref(h), 0, 1, 0, 0, 0, 0) == -1
    assert L.urgym_create(ctypes.byref(h), 0, 1, 4, 0, 0, 999) == -1
    assert L.urgym_create(ctypes.byref(h), 0, 1, 4, 0, 0, 0) == 0
    assert L.urgym_step(h, None, None, None, None, None, None, None, None, None, None, None) == -1
    assert b"must not be NULL" in L.urgym_last_error(h)
    assert L.urgym_get_state(h, 2, ctypes.c_void_p(1), None) == -5          # Ori has no obstacle
    assert L.urgym_destroy(h) == 0


def test_sb3_style_vec_env_contract(ug):
    """DummyVecEnv semantics train.py relies on: auto-reset, terminal_observation, TimeLimit.truncated, Monitor's
    episode info, observation dict of numpy arrays"""
    n = 512
    venv = ug.SB3VecEnvAdapter("UR5DynReach-v1", n, seed=3)
    ref = ug.UR5VecEnv("UR5DynReach-v1", n, seed=3)
    obs = venv.reset(); ref.reset()
    assert set(obs) == {"observation", "achieved_goal", "desired_goal"} and obs["observation"].shape == (n, 35)
    assert obs["observation"].dtype == np.float32 and venv.action_space.shape == (6,)
    rng = np.random.default_rng(0)
    ret, length = np.zeros(n), np.zeros(n, int)
    seen_done = seen_trunc = 0
    for t in range(130):
        a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        a[: n // 4] *= 0.02                                   # a quarter of the envs creep -> TimeLimit
        obs, rew, dones, infos = venv.step(a)
        o2, r2, te, tr, inf = ref.step(torch.from_numpy(a).cuda())
        assert np.array_equal(obs["observation"], o2["observation"].cpu().numpy())
        assert np.array_equal(dones, (te | tr).bool().cpu().numpy()) and len(infos) == n
        ret += rew; length += 1
        for i in np.nonzero(dones)[0]:
            info = infos[i]
            assert np.array_equal(info["terminal_observation"]["observation"], inf["terminal_observation"][i].cpu().numpy())
            assert abs(info["episode"]["r"] - ret[i]) < 1e-3 * max(1.0, abs(ret[i])) and info["episode"]["l"] == length[i]
            assert info["TimeLimit.truncated"] == bool(tr[i].item() and not te[i].item())
            seen_trunc += info["TimeLimit.truncated"]
            ret[i] = 0; length[i] = 0; seen_done += 1
        assert all("terminal_observation" not in infos[i] for i in np.nonzero(~dones)[0][:20])
    assert seen_done > 0 and seen_trunc > 0
    venv.close()


# ---------------------------------------------------------------------------------------------------- round 2
@pytest.mark.parametrize("env_id,n_total,shard", [("UR5DynReach-v1", 1 << 20, None), ("UR5ObsReach-v1", 1 << 20, None),
                                                  ("UR5OriReach-v1", 65536, None), ("UR5StaReach-v1", 1 << 21, 1)])
def test_full_size_oracle_sample(env_id, n_total, shard, ug):
    """Oracle parity AT the sizes BASELINE.json names (C5 per GPU, C3, C2, and the second 2 Mi shard of C4's 4 Mi envs):
    2 048 global env indices spread over the whole batch, one oracle env per index (the reset stream is keyed by the
    global index), 30 steps with auto-reset, every observation / reward / flag / terminal row / post-reset row compared
    at the north-star tolerances while the other envs of the batch step alongside."""
    from tests._gpu import GpuSampleSim
    offset = 0 if shard is None else shard * n_total
    rng = np.random.default_rng(17)
    k = 2048
    # whole tiles' first / last rows, the batch's first and last env, and a uniform sample
    fixed = [0, 1, 31, 32, 127, 128, n_total - 1, n_total - 2, n_total - 129]
    idx = np.unique(np.concatenate([fixed, rng.choice(n_total, k - len(fixed), replace=False)]))[:k] + offset
    sim = GpuSampleSim(env_id, n_total, idx, seed=31, offset=offset)
    st = run_parity(sim, env_id, oe.GEOM_CAPSULE, len(idx), 30, seed=31, indices=idx)
    assert st["steps"] > 0.9 * len(idx) * 30 and st["resets"] > 0, st
    print(env_id, n_total, st)


@pytest.mark.parametrize("env_id", ["UR5ObsReach-v1", "UR5StaReach-v1", "UR5DynReach-v1"])
def test_workbench_link_dist_mode_parity(env_id, ug):
    from tests._gpu import GpuSim
    n, steps = 200, 60
    st = run_parity(GpuSim(env_id, oe.GEOM_CAPSULE, n, seed=21, offset=300, link_dist_mode=oe.LD_WORKBENCH), env_id,
                    oe.GEOM_CAPSULE, n, steps, seed=21, offset=300, link_dist_mode=oe.LD_WORKBENCH)
    assert st["steps"] > 0.85 * n * steps and st["resets"] > 0, st
    # and the mode is what it says: never larger than the obstacle-only distance, smaller for links near the table
    a = ug.UR5VecEnv(env_id, 4096, seed=5, link_dist="workbench"); b = ug.UR5VecEnv(env_id, 4096, seed=5)
    a.reset(); b.reset()
    la, lb = a.get_state("link_dist"), b.get_state("link_dist")
    assert (la <= lb + 1e-7).all() and (la < lb - 1e-3).any()


@pytest.mark.parametrize("geom", [oe.GEOM_CAPSULE, oe.GEOM_HULL])
def test_sta_moving_obstacle_injection(geom, ug):
    """ReachSta's 18-value injection (reach.py:483-503) + set_velocity (reach.py:518-541) through urgym_set_state"""
    from tests._gpu import GpuSim
    n, steps = (200, 70) if geom == oe.GEOM_CAPSULE else (40, 35)
    sc = sta_moving_scenarios(n, seed=5)

    def inject(sim, orc):
        sim.set_goal(sc[:, :6]); sim.set_obstacle(sc[:, 6:12]); sim.set_obstacle_start(sc[:, 6:12]); sim.set_obstacle_end(sc[:, 12:])
        sim.refresh()
        for i, e in enumerate(orc.envs):
            e.task.set_goal_and_obstacle(sc[i].astype(np.float64))

    # (hull: a tilting cylinder brings its rims to the links more often than the static scenes do; FP32 GJK against a rim
    # converges sublinearly)
    kw = dict(ld_tol=HULL_LD_TOL, rew_atol=HULL_REW_ATOL) if geom == oe.GEOM_HULL else {}
    st = run_parity(GpuSim("UR5StaReach-v1", geom, n, seed=8, offset=40), "UR5StaReach-v1", geom, n, steps, seed=8, offset=40,
                    action_scale=0.6, after_reset=inject, **kw)
    assert st["steps"] > 0.8 * n * steps and st["resets"] > 0, st


def test_sta_moving_obstacle_through_make(ug):
    """the reference-facing surface: env.task.set_goal_and_obstacle(18 values) on make('UR5StaReach-v1')"""
    env = ug.make("UR5StaReach-v1", seed=0)
    orc = oe.make("UR5StaReach-v1", geom=oe.GEOM_CAPSULE, stream=oe.PhiloxStream(0), env_index=0, first_event=1)
    env.reset(); orc.reset(event=2)
    data = np.array([0.5, 0.1, 0.1, -2.0, 0.0, -1.0, 0.6, -0.4, 0.3, 1.0, 0.8, 0.0, 0.9, 0.4, 0.5, -2.0, -1.2, 0.0])
    env.task.set_goal_and_obstacle(data); orc.task.set_goal_and_obstacle(data)
    np.testing.assert_allclose(env.task.obstacle_end, data[12:], atol=1e-6)
    np.testing.assert_allclose(env.task.obstacle_start, data[6:12], atol=1e-6)
    rng = np.random.default_rng(1)
    moved = 0.0
    for t in range(40):
        a = rng.uniform(-0.2, 0.2, 6).astype(np.float32)
        o, r, term, trunc, info = env.step(a)
        o2, r2, term2, trunc2, info2 = orc.step(a)
        lin, ang = obs_close("UR5StaReach-v1", o["observation"], o2["observation"])
        assert lin <= 1e-5 and ang <= 1e-5, (t, lin, ang)
        assert abs(r - r2) <= 1e-5 * max(1.0, abs(r2)) and term == term2
        moved = max(moved, float(np.linalg.norm(o["observation"][18:21] - data[6:9])))
        if term:
            break
    assert moved > 0.5          # the obstacle travelled to (within 0.05 m of) its end point
    with pytest.raises(ValueError):
        env.task.set_goal_and_obstacle(np.zeros(7))
    env.close()


def test_use_before_reset_raises_and_checkpoint_is_validated(ug):
    env = ug.UR5VecEnv("UR5DynReach-v1", 64, seed=1)
    with pytest.raises(ug.UrgymError):
        env.step(torch.zeros((64, 6), device="cuda"))
    with pytest.raises(ug.UrgymError):
        env.observe()
    env.reset()
    ck = env.state_dict()
    other = ug.UR5VecEnv("UR5DynReach-v1", 64, seed=2)
    with pytest.raises(ug.UrgymError):
        other.load_state_dict(ck)                       # another seed: the reset stream would differ
    shuffled = dict(reversed(list(ck.items())))         # the dict's own order does not matter
    same = ug.UR5VecEnv("UR5DynReach-v1", 64, seed=1)
    same.load_state_dict(shuffled)
    a = torch.rand((64, 6), device="cuda") * 2 - 1
    assert torch.equal(env.step(a)[0]["observation"], same.step(a)[0]["observation"])


def test_chain_layout_change_never_reuses_reset_stream_positions(ug):
    """2-chain graph, then a 4-chain graph of the same handle: urgym_sync_events at the head of every chained graph
    raises the lagging chains' counters, so the second layout draws fresh episodes."""
    env_id, n, k = "UR5DynReach-v1", 40_000, 5
    env = ug.UR5VecEnv(env_id, n, seed=3); env.reset()
    g = torch.Generator(device="cuda").manual_seed(2)
    ring = [torch.rand((n, 6), device="cuda", generator=g) * 2.4 - 1.2 for _ in range(k)]
    g2 = env.capture_steps(ring, chains=2)
    for _ in range(3):
        g2.replay()
    torch.cuda.synchronize()
    ev = ctypes.c_uint32(); env.L.urgym_get_event(env.h, ctypes.byref(ev))
    assert ev.value == 1 + 3 * k
    g4 = env.capture_steps(ring, chains=4)
    g4.replay(); torch.cuda.synchronize()
    env.L.urgym_get_event(env.h, ctypes.byref(ev))
    assert ev.value == 1 + 4 * k
    # reference run: the same 20 steps as whole-batch steps
    ref = ug.UR5VecEnv(env_id, n, seed=3); ref.reset()
    for _ in range(4):
        for a in ring:
            ref.step(a)
    assert torch.equal(ref.obs, env.obs) and torch.equal(ref.get_state("goal"), env.get_state("goal"))


def _shipped_policy(ug, env_id, device):
    from tests.closed_loop import SHORT
    import os
    w = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"policy_{SHORT[env_id]}.npz"))
    return ug.mlp_policy({k: torch.as_tensor(w[k], device=device) for k in w.files if not k.startswith("published")})


def test_device_rollout_matches_vecenv_transitions(ug):
    """Rollout (policy -> urgym_step -> urgym_replay_write, all on the device) stores exactly the transitions an SB3-style
    loop over SB3VecEnvAdapter would put into its replay buffer (train.py:39-60 through DummyVecEnv): obs, action, reward,
    next_obs = terminal_observation for finished envs, done, TimeLimit.truncated."""
    env_id, n, steps = "UR5DynReach-v1", 384, 45
    env = ug.UR5VecEnv(env_id, n, seed=6)
    env.reset()
    policy = _shipped_policy(ug, env_id, env.device)
    noisy = lambda o: torch.clamp(policy(o) + 0.6 * torch.sin(o["observation"][:, :6] * 37.0), -1, 1)   # exploration, deterministic
    ring = ug.DeviceReplayRing(env, capacity=n * steps)
    ro = ug.Rollout(env, noisy, ring)
    ro.run(steps)
    torch.cuda.synchronize()
    assert int(ring.cursor.item()) == n * steps and len(ring) == n * steps
    venv = ug.SB3VecEnvAdapter(env_id, n, seed=6)
    obs = venv.reset()
    seen_done = seen_timeout = 0
    for t in range(steps):
        od = {k: torch.as_tensor(v).cuda() for k, v in obs.items()}
        a = noisy(od).cpu().numpy()
        nobs, rew, dones, infos = venv.step(a)
        rows = slice(t * n, (t + 1) * n)
        assert np.array_equal(ring.obs[rows].cpu().numpy(), obs["observation"])
        assert np.array_equal(ring.actions[rows].cpu().numpy(), a)
        assert np.array_equal(ring.reward[rows].cpu().numpy(), rew)
        assert np.array_equal(ring.done[rows].cpu().numpy().astype(bool), dones)
        nxt = nobs["observation"].copy()
        for i in np.nonzero(dones)[0]:
            nxt[i] = infos[i]["terminal_observation"]["observation"]
        assert np.array_equal(ring.next_obs[rows].cpu().numpy(), nxt)
        to = np.array([infos[i]["TimeLimit.truncated"] for i in range(n)])
        assert np.array_equal(ring.timeout[rows].cpu().numpy().astype(bool), to)
        seen_done += int(dones.sum()); seen_timeout += int(to.sum())
        obs = nobs
    assert seen_done > 0
    batch = ring.sample(256)
    assert batch["observations"].shape == (256, 35) and batch["dones"].max() <= 1.0
    venv.close()


def test_device_rollout_graph_equals_eager_and_ring_wraps(ug):
    env_id, n = "UR5StaReach-v1", 2048
    envs = [ug.UR5VecEnv(env_id, n, seed=4) for _ in range(2)]
    for e in envs:
        e.reset()
    policy = _shipped_policy(ug, env_id, envs[0].device)
    rings = [ug.DeviceReplayRing(e, capacity=n * 6) for e in envs]        # 6 steps of capacity: the ring wraps
    eager, graphed = ug.Rollout(envs[0], policy, rings[0]), ug.Rollout(envs[1], policy, rings[1])
    graphed.capture(4)            # (one eager warm-up step inside capture())
    graphed.replay(3)
    eager.run(1 + 4 * 3)
    torch.cuda.synchronize()
    for name in ("obs", "next_obs", "actions", "reward", "done", "timeout", "cursor"):
        assert torch.equal(getattr(rings[0], name), getattr(rings[1], name)), name
    assert int(rings[0].cursor.item()) == 13 * n and len(rings[0]) == 6 * n
    assert torch.equal(envs[0].obs, envs[1].obs)


def test_async_host_step_matches_sync(ug):
    """urgym_step_host_async with two slots in flight == urgym_step_host, step for step"""
    env_id, n = "UR5DynReach-v1", 300_000
    a_env, b_env = ug.UR5VecEnv(env_id, n, seed=12), ug.UR5VecEnv(env_id, n, seed=12)
    a_env.reset(); b_env.reset()
    sync = a_env.alloc_host_buffers()
    slots = [b_env.alloc_host_buffers() for _ in range(2)]
    rng = np.random.default_rng(3)
    acts = [torch.from_numpy(rng.uniform(-1, 1, (n, 6)).astype(np.float32)).pin_memory() for _ in range(7)]
    expect = []
    for a in acts:
        a_env.step_host(a, sync)
        expect.append({k: v.clone() for k, v in sync.items() if k != "actions"})
    b_env.step_host_async(0, acts[0], slots[0])
    for k in range(1, len(acts)):
        b_env.step_host_async(k % 2, acts[k], slots[k % 2])      # enqueued before the previous step was waited for
        b_env.host_wait((k - 1) % 2)
        done = (expect[k - 1]["terminated"] | expect[k - 1]["truncated"]).bool()
        for key in ("obs", "reward", "terminated", "truncated", "is_success"):
            assert torch.equal(slots[(k - 1) % 2][key], expect[k - 1][key]), (k, key)
        assert torch.equal(slots[(k - 1) % 2]["terminal_obs"][done], expect[k - 1]["terminal_obs"][done])
    b_env.host_wait((len(acts) - 1) % 2)
    assert torch.equal(slots[(len(acts) - 1) % 2]["obs"], expect[-1]["obs"])
    with pytest.raises(ug.UrgymError):
        b_env.step_host_async(0, acts[0], slots[0]); b_env.step_host_async(0, acts[1], slots[0])   # slot still in flight
    b_env.host_wait(0)
