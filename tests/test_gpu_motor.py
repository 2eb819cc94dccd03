"""GPU tier: the motor-driven env UR5IAIReach-v1 (urgym_motor_* entry points) against its CPU oracle
(oracle/motor_oracle.py + oracle/ur_motor_oracle.c) on the same seeds and actions.  Tolerances are the north-star's:
50-step joint trajectories within 1e-3 rad, positions within 1e-5 m per step from identical injected states, rewards 1e-5
relative, flags exact (1e-6 band).  Both sides restate Bullet (PARITY UNPINNED): this checks FP32 link-frame recursive
Newton-Euler on the GPU against FP64 world-frame Jacobians on the CPU, not either of them against PyBullet."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _vec(n, seed=0, offset=0):
    import urgym_b200 as ug
    return ug.UR5MotorVecEnv(n, device=0, seed=seed, env_index_offset=offset)


def _oracles(n, seed=0, offset=0):
    from oracle import motor_oracle as mo
    from oracle import oracle_env as oe
    return [mo.UR5IAIReachOracle(oe.PhiloxStream(seed), env_index=offset + i, first_event=1) for i in range(n)]


def test_rollout_parity_50_steps():
    n, steps = 48, 50
    vec, orc = _vec(n, seed=11, offset=1000), _oracles(n, seed=11, offset=1000)
    o = vec.reset()
    torch.cuda.synchronize()
    ref0 = np.stack([e._get_obs()["observation"] for e in orc])
    assert np.abs(o["observation"].cpu().numpy() - ref0).max() < 1e-6
    assert np.abs(o["desired_goal"].cpu().numpy() - np.stack([e.goal for e in orc])).max() < 1e-6
    rng = np.random.default_rng(5)
    worst_q = worst_p = worst_v = worst_r = 0.0
    for k in range(steps):
        a = rng.uniform(-1.2, 1.2, (n, 6)).astype(np.float32)
        obs, rew, term, trunc, info = vec.step(torch.as_tensor(a).cuda())
        torch.cuda.synchronize()
        q = vec.get_state("q").cpu().numpy()
        for i, e in enumerate(orc):
            eo, er, et, etr, _ = e.step(a[i])
            assert bool(term[i].item()) == et and bool(trunc[i].item()) == etr
            worst_q = max(worst_q, np.abs(q[i] - e.q).max())
            worst_p = max(worst_p, np.abs(obs["observation"][i, :3].cpu().numpy() - eo["observation"][:3]).max())
            worst_v = max(worst_v, np.abs(obs["observation"][i, 3:].cpu().numpy() - eo["observation"][3:]).max())
            worst_r = max(worst_r, abs(float(rew[i].item()) - er) / max(1.0, abs(er)))
    print(f"motor parity over {steps} steps: joints {worst_q:.2e} rad, ee position {worst_p:.2e} m, ee velocity {worst_v:.2e} m/s, "
          f"reward {worst_r:.2e}")
    assert worst_q < 1e-3                       # north-star: 50-step joint trajectories within 1e-3 rad
    assert worst_p < 1e-3 and worst_v < 2e-2 and worst_r < 1e-3      # accumulated over the trajectory


def test_single_step_from_injected_states():
    """identical injected joint states / velocities / goals, one step: the per-step tolerances"""
    from oracle import motor_oracle as mo
    n = 256
    rng = np.random.default_rng(9)
    q0 = rng.uniform(-2.0, 2.0, (n, 6)); qd0 = rng.uniform(-1.0, 1.0, (n, 6)); goal = rng.uniform(mo.GOAL_LOW, mo.GOAL_HIGH, (n, 3))
    a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
    vec = _vec(n)
    vec.reset()
    vec.set_state("q", q0.astype(np.float32)); vec.set_state("qd", qd0.astype(np.float32)); vec.set_state("goal", goal.astype(np.float32))
    obs, rew, term, trunc, info = vec.step(torch.as_tensor(a).cuda())
    torch.cuda.synchronize()
    q1 = vec.get_state("q").cpu().numpy(); qd1 = vec.get_state("qd").cpu().numpy()
    wq = wv = wp = wr = 0.0
    for i in range(n):
        target = q0[i].astype(np.float32).astype(np.float64) + ((np.clip(a[i], -1, 1) * np.float32(np.pi)) * np.float32(0.1)).astype(np.float64)
        qr, qdr = mo.substeps(q0[i].astype(np.float32), qd0[i].astype(np.float32), target)
        pos, vel = mo.ee_state(qr, qdr)
        d = float(np.linalg.norm(pos - goal[i].astype(np.float32)))
        wq = max(wq, np.abs(q1[i] - qr).max()); wv = max(wv, np.abs(qd1[i] - qdr).max())
        wp = max(wp, np.abs(obs["observation"][i, :3].cpu().numpy() - pos).max())
        wr = max(wr, abs(float(rew[i].item()) + d) / max(1.0, d))
    print(f"motor single step: joints {wq:.2e} rad, joint velocities {wv:.2e} rad/s, ee position {wp:.2e} m, reward {wr:.2e}")
    assert wq < 1e-5 and wp < 1e-5 and wr < 1e-5 and wv < 5e-4


def test_timelimit_autoreset_and_terminal_rows():
    n = 40
    vec, orc = _vec(n, seed=3), _oracles(n, seed=3)
    vec.reset()
    zero = torch.zeros((n, 6), device="cuda")
    for k in range(100):
        obs, rew, term, trunc, info = vec.step(zero)
    torch.cuda.synchronize()
    assert trunc.cpu().numpy().all() and not term.cpu().numpy().any()
    # the envs restarted inside the call: neutral pose at rest, a new goal from reset event 102 (1 reset + 100 steps, then this one)
    assert np.abs(vec.get_state("q").cpu().numpy() - np.array([0.0, -1.5708, 0.0, 0.0, 0.0, 0.0], np.float32)).max() == 0.0
    assert np.abs(vec.get_state("qd").cpu().numpy()).max() == 0.0
    assert (vec.get_state("elapsed").cpu().numpy() == 0).all()
    for i, e in enumerate(orc):
        e.reset(event=101)
    assert np.abs(obs["desired_goal"].cpu().numpy() - np.stack([e.goal for e in orc])).max() < 1e-6
    # terminal rows = the holding arm's last observation (it sags < 1 mm under gravity in 4 s), not the reset observation
    t = info["terminal_observation"].cpu().numpy()
    assert np.abs(t[:, :3] - obs["observation"].cpu().numpy()[:, :3]).max() < 2e-3
    st = vec.stats()
    assert st["episodes"] == n and st["truncations"] == n and st["length_sum"] == 100 * n


def test_make_surface_and_errors():
    import urgym_b200 as ug
    env = ug.make("UR5IAIReach-v1", render=False)
    o, info = env.reset()
    assert o["observation"].shape == (6,) and o["achieved_goal"].shape == (3,) and o["desired_goal"].shape == (3,)
    env.task.set_goal(np.array([0.4, 0.1, 0.5]))
    env.robot.set_joint_angles(np.array([0.1, -1.2, 0.3, 0.0, 0.2, 0.0]))
    o, r, term, trunc, info = env.step(np.zeros(6, np.float32))
    assert r == pytest.approx(-float(np.linalg.norm(o["achieved_goal"] - np.array([0.4, 0.1, 0.5], np.float32))), abs=1e-5)
    assert not term and not trunc and info == {"is_success": False}
    env.close()
    L = ug._native.lib()
    h = ctypes.c_void_p()
    assert L.urgym_motor_create(ctypes.byref(h), 0, 0, ctypes.c_uint64(0), 0) == -1          # URGYM_EINVAL
    assert L.urgym_motor_create(ctypes.byref(h), 8, 0, ctypes.c_uint64(0), 99) == -2         # URGYM_ENODEVICE
    assert b"no usable CUDA device" in L.urgym_motor_last_error(None)
    assert L.urgym_motor_create(ctypes.byref(h), 8, 0, ctypes.c_uint64(0), 0) == 0
    assert L.urgym_motor_step(h, None, None, None, None, None, None, None, None, None, None) == -1
    assert L.urgym_motor_get_state(h, 7, ctypes.c_void_p(1), None) == -1
    assert L.urgym_motor_destroy(h) == 0
