"""The reference's motor-driven env `UR5IAIReach-v1` (UR_gym/envs/ur_tasks.py:10-21: robot `UR5`, UR_gym/envs/robots/UR5.py:10-118;
task `ReachIAI`, UR_gym/envs/tasks/reach.py:9-66) on the GPU: set_action -> POSITION_CONTROL motors -> 20 dynamic substeps
(pyb_setup.py:52-55,365-380) instead of the joint teleport of the four UR5e tasks.  SURVEY.md 8 f-4.

`UR5MotorVecEnv` is the batched API (DummyVecEnv semantics like UR5VecEnv); `MotorTaskEnv` is the single-env surface
`make("UR5IAIReach-v1")` returns.  Bullet's substep is restated, not pinned to a PyBullet run (include/urgym_b200.h)."""
import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from . import _native as nat

OBS_DIM, GOAL_DIM = 6, 3
GOAL_RANGE_LOW = np.array([0.2, -0.4, 0.0])        # reach.py:20
GOAL_RANGE_HIGH = np.array([0.6, 0.4, 0.8])        # reach.py:21


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class UR5MotorVecEnv:
    """num_envs independent `UR5IAIReach-v1` envs on one GPU.  step(actions [N,6]) -> (obs dict, reward, terminated,
    truncated, info); finished envs restart inside the same call, info["terminal_observation"] keeps their last rows."""

    def __init__(self, num_envs: int, device: int = 0, seed: int = 0, env_index_offset: int = 0):
        if not torch.cuda.is_available():
            raise nat.UrgymError("no CUDA device: ur-gym_b200 has no CPU path")
        self.env_id = nat.MOTOR_ENV_ID
        self.num_envs, self.device_index = int(num_envs), int(device)
        self.device = torch.device("cuda", self.device_index)
        self.L = nat.lib()
        h = ctypes.c_void_p()
        nat.check_motor(None, self.L.urgym_motor_create(ctypes.byref(h), self.num_envs, int(env_index_offset),
                                                        ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), self.device_index))
        self.h = h
        n, kw = self.num_envs, dict(device=self.device)
        self.obs = torch.zeros((n, OBS_DIM), dtype=torch.float32, **kw)
        self.achieved = torch.zeros((n, GOAL_DIM), dtype=torch.float32, **kw)
        self.desired = torch.zeros((n, GOAL_DIM), dtype=torch.float32, **kw)
        self.terminal_obs = torch.zeros((n, OBS_DIM), dtype=torch.float32, **kw)
        self.reward = torch.zeros(n, dtype=torch.float32, **kw)
        self.terminated = torch.zeros(n, dtype=torch.uint8, **kw)
        self.truncated = torch.zeros(n, dtype=torch.uint8, **kw)
        self.is_success = torch.zeros(n, dtype=torch.uint8, **kw)
        self._obs = {"observation": self.obs, "achieved_goal": self.achieved, "desired_goal": self.desired}
        self._info = {"is_success": self.is_success, "terminal_observation": self.terminal_obs}
        self._initialised = False

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None):
            self.L.urgym_motor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, mask: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        if mask is not None:
            if not self._initialised:
                raise nat.UrgymError("reset() of all envs must come first")
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        self._initialised = True
        nat.check_motor(self.h, self.L.urgym_motor_reset(self.h, _ptr(mask), _ptr(self.obs), _ptr(self.achieved),
                                                         _ptr(self.desired), self._stream()))
        return self._obs

    def step(self, actions: torch.Tensor):
        if not self._initialised:
            raise nat.UrgymError("reset() must be called before step()")
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
        if actions.shape != (self.num_envs, 6):
            raise ValueError(f"actions must have shape ({self.num_envs}, 6)")
        nat.check_motor(self.h, self.L.urgym_motor_step(self.h, _ptr(actions), _ptr(self.obs), _ptr(self.achieved),
                                                        _ptr(self.desired), _ptr(self.reward), _ptr(self.terminated),
                                                        _ptr(self.truncated), _ptr(self.is_success), _ptr(self.terminal_obs),
                                                        self._stream()))
        return self._obs, self.reward, self.terminated, self.truncated, self._info

    _FIELDS = {"q": (nat.MOTOR_F_Q, 6, torch.float32), "qd": (nat.MOTOR_F_QD, 6, torch.float32),
               "goal": (nat.MOTOR_F_GOAL, 3, torch.float32), "elapsed": (nat.MOTOR_F_ELAPSED, 0, torch.int32)}

    def get_state(self, name: str) -> torch.Tensor:
        f, w, dt = self._FIELDS[name]
        out = torch.empty((self.num_envs, w) if w else (self.num_envs,), dtype=dt, device=self.device)
        nat.check_motor(self.h, self.L.urgym_motor_get_state(self.h, f, _ptr(out), self._stream()))
        return out

    def set_state(self, name: str, value) -> None:
        f, w, dt = self._FIELDS[name]
        v = torch.as_tensor(value, dtype=dt, device=self.device).reshape((self.num_envs, w) if w else (self.num_envs,)).contiguous()
        nat.check_motor(self.h, self.L.urgym_motor_set_state(self.h, f, _ptr(v), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()      # `v` may be a temporary

    def stats(self, reset: bool = True) -> Dict[str, float]:
        torch.cuda.synchronize(self.device)
        out = (ctypes.c_double * 8)()
        nat.check_motor(self.h, self.L.urgym_motor_stats(self.h, out, 1 if reset else 0))
        return dict(zip(nat.STAT_NAMES, list(out)))


class _MotorRobot:
    """UR5 (UR5.py:10-118) as far as callers touch it"""

    def __init__(self, env):
        self._e = env
        self.neutral_joint_values = np.array([0.0, -1.5708, 0.0, 0.0, 0.0, 0.0])

    def set_joint_angles(self, angles) -> None:          # core.py:161-167: resetJointState (velocities zeroed)
        self._e.vec.set_state("q", np.asarray(angles, np.float32).reshape(1, 6))
        self._e.vec.set_state("qd", np.zeros((1, 6), np.float32))

    def get_joint_angles(self) -> np.ndarray:
        return self._e.vec.get_state("q")[0].cpu().numpy().astype(np.float64)


class _MotorTask:
    """ReachIAI (reach.py:9-66)"""
    distance_threshold = 0.005
    goal_range_low, goal_range_high = GOAL_RANGE_LOW, GOAL_RANGE_HIGH

    def __init__(self, env):
        self._e = env

    def get_goal(self) -> np.ndarray:
        return self._e.vec.get_state("goal")[0].cpu().numpy().astype(np.float64)

    def set_goal(self, goal) -> None:
        self._e.vec.set_state("goal", np.asarray(goal, np.float32).reshape(1, 3))

    def is_success(self, achieved_goal, desired_goal):
        d = np.sqrt(((np.asarray(achieved_goal)[..., :3] - np.asarray(desired_goal)[..., :3]) ** 2).sum(-1))
        return np.array(d < self.distance_threshold)

    def compute_reward(self, achieved_goal, desired_goal, info=None):
        d = np.sqrt(((np.asarray(achieved_goal)[..., :3] - np.asarray(desired_goal)[..., :3]) ** 2).sum(-1))
        return -d.astype(np.float32)


class MotorTaskEnv:
    """gymnasium.make("UR5IAIReach-v1", render=False): reset(seed, options) -> (obs, info); step(a) -> (obs, reward,
    terminated, truncated, info).  One env of a UR5MotorVecEnv without auto-reset-on-done semantics exposed: like
    gymnasium's TimeLimit wrapper the env must be reset by the caller after terminated / truncated."""

    def __init__(self, env_id: str = nat.MOTOR_ENV_ID, render: bool = False, device: int = 0, seed: int = 0):
        if env_id != nat.MOTOR_ENV_ID:
            raise ValueError(f"unknown motor env id {env_id!r}")
        if render:
            raise nat.UrgymError("rendering is out of scope (PyBullet GUI); use render=False")
        self.vec = UR5MotorVecEnv(1, device=device, seed=seed)
        self.robot, self.task = _MotorRobot(self), _MotorTask(self)
        self.compute_reward = self.task.compute_reward
        self.reset()

    def _np(self, obs):
        return {k: v[0].cpu().numpy().copy() for k, v in obs.items()}

    def reset(self, seed=None, options=None):
        o = self._np(self.vec.reset())
        self._goal_before = o["desired_goal"]
        return o, {"is_success": False}

    def step(self, action):
        a = torch.as_tensor(np.asarray(action, np.float32).reshape(1, 6), device=self.vec.device)
        obs, r, term, trunc, info = self.vec.step(a)
        term_b, trunc_b = bool(term[0].item()), bool(trunc[0].item())
        if term_b or trunc_b:       # the batch API has already restarted the env: hand out the episode's own last rows
            o = {"observation": info["terminal_observation"][0].cpu().numpy().copy(),
                 "achieved_goal": info["terminal_observation"][0, :3].cpu().numpy().copy(),
                 "desired_goal": self._goal_before}
        else:
            o = self._np(obs)
        self._goal_before = o["desired_goal"]
        return o, float(r[0].item()), term_b, trunc_b, {"is_success": bool(info["is_success"][0].item())}

    def close(self):
        self.vec.close()
