"""gymnasium-style single-environment surface of the reference, backed by the CUDA simulator with N = 1:

    env = make("UR5DynReach-v1", render=False); obs, info = env.reset(); obs, r, term, trunc, info = env.step(a)

Mirrors UR_gym/__init__.py:19-42 (ids, max_episode_steps=100), UR_gym/envs/ur_tasks.py:37-90 and
UR_gym/envs/core.py:222-317 (RobotTaskEnv), including the attribute access the reference's own scripts use:
env.task.set_goal / set_goal_and_obstacle / goal_range_low / get_obs / get_goal / is_success, env.robot.get_obs,
env.compute_reward (model_test.py:11-38, utils/generate.py:30-31).  Everything numeric comes from the GPU library;
this file is plumbing."""
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from . import _native as nat
from .utils import angular_distance, distance
from .vec_env import UR5VecEnv

try:                                    # gymnasium is optional (absent in the build image)
    from gymnasium import spaces as _spaces
    Box, DictSpace = _spaces.Box, _spaces.Dict
except Exception:                        # minimal stand-ins with the attributes callers read
    class Box:
        def __init__(self, low, high, shape, dtype=np.float32):
            self.low = np.full(shape, low, dtype=dtype); self.high = np.full(shape, high, dtype=dtype)
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class DictSpace(dict):
        def __init__(self, d):
            super().__init__(d); self.spaces = self

        def sample(self):
            return {k: v.sample() for k, v in self.items()}

_GOAL_RANGES = {   # reach.py:151-152, 248-251, 385-388, 584-587
    "UR5OriReach-v1": ([0.3, -0.5, 0.0], [0.75, 0.5, 0.2], None, None),
    "UR5ObsReach-v1": ([0.3, -0.5, -0.1], [0.75, 0.5, 0.2], [0.5, -0.5, 0.25], [1.0, 0.5, 0.55]),
    "UR5StaReach-v1": ([0.3, -0.5, 0.0], [0.75, 0.5, 0.2], [0.5, -0.5, 0.25], [1.0, 0.5, 0.55]),
    "UR5DynReach-v1": ([0.4, -0.5, 0.0], [0.75, 0.5, 0.2], [0.5, -0.8, 0.25], [1.2, 0.8, 0.75]),
}


class _Spec:
    def __init__(self, env_id):
        self.id, self.max_episode_steps = env_id, 100


class Robot:
    """UR5Ori surface (UR_gym/envs/robots/UR5.py:243-351)."""

    def __init__(self, env):
        self._env = env
        self.action_space = Box(-1.0, 1.0, shape=(6,), dtype=np.float32)        # UR5.py:251
        self.neutral_joint_values = np.array([0.0, -1.5708, 0.0, -1.5708, 0.0, 0.0])   # UR5.py:262

    def get_obs(self) -> np.ndarray:                       # UR5.py:320-325
        return self._env._observe()[:12].astype(np.float64)

    def get_joint_angles(self) -> np.ndarray:              # UR5.py:346-351
        return self._env.vec.get_state("q")[0].cpu().numpy().astype(np.float64)

    def set_joint_angles(self, angles) -> None:            # core.py:161-167
        self._env.vec.set_state("q", np.asarray(angles, np.float32)[None])

    def get_ee_position(self) -> np.ndarray:               # UR5.py:334-336
        return self.get_obs()[:3]

    def get_ee_orientation(self) -> np.ndarray:            # UR5.py:338-340
        return self.get_obs()[3:6]


class Task:
    """Reach{Ori,Obs,Sta,Dyn} surface (UR_gym/envs/tasks/reach.py)."""

    def __init__(self, env):
        self._env = env
        lo, hi, olo, ohi = _GOAL_RANGES[env.spec.id]
        self.goal_range_low, self.goal_range_high = np.array(lo), np.array(hi)
        if olo is not None:
            self.obs_range_low, self.obs_range_high = np.array(olo), np.array(ohi)
        self.distance_threshold, self.ori_distance_threshold = 0.05, 0.0873
        self.collision = False

    # -- read-outs
    @property
    def goal(self) -> np.ndarray:
        return self._env.vec.get_state("goal")[0].cpu().numpy().astype(np.float64)

    def get_goal(self) -> np.ndarray:                      # core.py:206-211
        return self.goal.copy()

    @property
    def link_dist(self) -> np.ndarray:
        return self._env.vec.get_state("link_dist")[0].cpu().numpy().astype(np.float64)

    @property
    def obstacle(self) -> np.ndarray:
        return self._env.vec.get_state("obstacle")[0].cpu().numpy().astype(np.float64)

    @property
    def obstacle_start(self) -> np.ndarray:
        return self._env.vec.get_state("obstacle_start")[0].cpu().numpy().astype(np.float64)

    @property
    def obstacle_end(self) -> np.ndarray:
        return self._env.vec.get_state("obstacle_end")[0].cpu().numpy().astype(np.float64)

    def get_obs(self) -> np.ndarray:                       # reach.py:189-190,307-308,454-458,653-657
        return self._env._observe()[12:].astype(np.float64)

    def get_achieved_goal(self) -> np.ndarray:             # reach.py:192-195,310-311,460-463,659-662
        return self._env._observe()[:self._env.vec.goal_dim].astype(np.float64)

    # -- injection hooks
    def set_goal(self, test_goal) -> None:                 # reach.py:202-204
        self._env.vec.set_state("goal", np.asarray(test_goal, np.float32)[None])

    def set_goal_and_obstacle(self, test_data) -> None:    # reach.py:328-335, 483-503, 702-713
        t = np.asarray(test_data, np.float32)
        v, G = self._env.vec, self._env.vec.goal_dim
        if self._env.spec.id == "UR5OriReach-v1":
            raise AttributeError("ReachOri has no set_goal_and_obstacle")
        if self._env.spec.id == "UR5DynReach-v1":
            if t.size != 18:
                raise ValueError("expected goal(6) + obstacle_start(6) + obstacle_end(6)")
            v.set_state("goal", t[None, :6]); v.set_state("obstacle", t[None, 6:12]); v.set_state("obstacle_end", t[None, 12:])
        elif self._env.spec.id == "UR5StaReach-v1" and t.size == 18:
            # moving obstacle: obstacle = obstacle_start = t[6:12], obstacle_end = t[12:]   reach.py:491-499
            v.set_state("goal", t[None, :6]); v.set_state("obstacle", t[None, 6:12])
            v.set_state("obstacle_start", t[None, 6:12]); v.set_state("obstacle_end", t[None, 12:])
        else:
            if t.size != G + 6:
                raise ValueError(f"expected {G + 6} values (goal + obstacle)" + (" or 18" if G == 6 else ""))
            v.set_state("goal", t[None, :G]); v.set_state("obstacle", t[None, G:])     # (Sta: start / end stay as they are)
        self.collision = bool(v.refresh()[0].item())

    # -- pure functions of arrays (utils.py formulas)
    def is_success(self, achieved_goal, desired_goal) -> np.ndarray:     # reach.py:212-215,348-350,543-546,755-758
        ok = distance(achieved_goal, desired_goal) < self.distance_threshold
        if self._env.spec.id != "UR5ObsReach-v1":
            ok = ok & (angular_distance(achieved_goal, desired_goal) < self.ori_distance_threshold)
        return np.array(ok, dtype=np.bool_)

    def compute_reward(self, achieved_goal, desired_goal, info=None) -> np.ndarray:
        """reach.py:221-236,356-374,552-573,764-785, evaluated with the collision flag and the link distances of the
        last step (read-only: unlike the reference it does not advance last_dist when called)."""
        e = self._env
        succ = self.is_success(achieved_goal, desired_goal)
        d = distance(achieved_goal, desired_goal)
        kind = e.spec.id
        if kind == "UR5OriReach-v1":
            return 200.0 * succ - 70.0 * d - 30.0 * angular_distance(achieved_goal, desired_goal) - 500.0 * self.collision
        change = np.where(e._ld_new < 0.2, e._ld_new - e._ld_prev, 0.0)
        if kind == "UR5ObsReach-v1":
            return 200.0 * succ - 500.0 * self.collision - 100.0 * d + (100.0 * change).sum()
        if self.collision:
            return np.float64(-500.0)
        if bool(np.all(succ)):
            return np.float64(200.0)
        w = np.array([8, 2.4, 1.2, 1.2, 0.2]) / 13.0 * 50.0
        return -70.0 * d - 30.0 * angular_distance(achieved_goal, desired_goal) + (w * change).sum()


class RobotTaskEnv:
    """One environment behind the reference's Env protocol (core.py:222-317) + TimeLimit(100)."""

    metadata = {"render_modes": ["human", "rgb_array"]}

    def __init__(self, env_id: str, render: bool = False, device: int = 0, seed: Optional[int] = None,
                 geometry: str = "capsule", env_index: int = 0, link_dist: str = "obstacle"):
        if render:
            raise NotImplementedError("the batched GPU simulator has no GUI; use render=False")
        self.spec = _Spec(env_id)
        self.vec = UR5VecEnv(env_id, 1, device=device, seed=0 if seed is None else seed, env_index_offset=env_index,
                             geometry=geometry, auto_reset=False, link_dist=link_dist)
        self.robot, self.task = Robot(self), Task(self)
        self.sim = self.vec
        self._ld_prev = self._ld_new = np.zeros(5)
        obs, _ = self.reset()                                          # core.py:237
        D, G = obs["observation"].shape, obs["achieved_goal"].shape
        self.observation_space = DictSpace(dict(observation=Box(-10.0, 10.0, shape=D, dtype=np.float32),
                                                desired_goal=Box(-10.0, 10.0, shape=G, dtype=np.float32),
                                                achieved_goal=Box(-10.0, 10.0, shape=G, dtype=np.float32)))
        self.action_space = self.robot.action_space
        self.compute_reward = self.task.compute_reward                 # core.py:249

    def _observe(self) -> np.ndarray:
        return self.vec.observe()["observation"][0].cpu().numpy()

    @staticmethod
    def _np(d: Dict[str, torch.Tensor]) -> Dict[str, np.ndarray]:
        return {k: v[0].cpu().numpy().copy() for k, v in d.items()}

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None) -> Tuple[Dict[str, np.ndarray], Dict[str, Any]]:
        if seed is not None:
            self.vec.reseed(seed)
        obs = self._np(self.vec.reset())
        self.task.collision = False
        if self.spec.id != "UR5OriReach-v1":
            self._ld_prev = self._ld_new = self.task.link_dist
        info = {"is_success": self.task.is_success(obs["achieved_goal"], obs["desired_goal"])}
        return obs, info

    def step(self, action) -> Tuple[Dict[str, np.ndarray], float, bool, bool, Dict[str, Any]]:
        a = torch.as_tensor(np.asarray(action, np.float32).reshape(1, 6))
        if self.spec.id != "UR5OriReach-v1":
            self._ld_prev = self.task.link_dist
        obs, rew, term, trunc, info = self.vec.step(a)
        obs = self._np(obs)
        terminated, truncated, success = bool(term[0].item()), bool(trunc[0].item()), bool(info["is_success"][0].item())
        self.task.collision = terminated and not success
        if self.spec.id != "UR5OriReach-v1":
            self._ld_new = self.task.link_dist
        return obs, float(rew[0].item()), terminated, truncated, {"is_success": success}

    def render(self):
        return None

    def close(self) -> None:
        self.vec.close()


ENV_IDS = tuple(nat.TASK_IDS)


def make(env_id: str, render: bool = False, **kwargs) -> RobotTaskEnv:
    """gymnasium.make(id, render=...) for the four reach tasks (UR_gym/__init__.py:19-42)."""
    if env_id == nat.MOTOR_ENV_ID:           # the motor-driven robot path (UR_gym/__init__.py:7-11, ur_tasks.py:10-21)
        from .motor_env import MotorTaskEnv
        return MotorTaskEnv(env_id, render=render, **kwargs)
    if env_id == "UR5RegReach-v1":
        # The reference registers this id (UR_gym/__init__.py:13-17) but its env cannot take a step: its scene has five
        # bodies (UR5, plane, table, track, target; reach.py:92-103) and PyBullet.check_collision reads keys[5]
        # (pyb_setup.py:398-399) -> IndexError at the first env.step().  There is no behaviour to reproduce.
        raise NotImplementedError("UR5RegReach-v1: the reference's own env raises IndexError at its first step "
                                  "(pyb_setup.py:398-399 reads keys[5] of a five-body scene); not built, see DESIGN.md section 8")
    if env_id not in nat.TASK_IDS:
        raise ValueError(f"unknown env id {env_id!r}; known: {list(ENV_IDS) + [nat.MOTOR_ENV_ID]}")
    return RobotTaskEnv(env_id, render=render, **kwargs)


def register_with_gymnasium() -> bool:
    """register the ids with gymnasium when it is installed (it is not in the build image)"""
    try:
        from gymnasium.envs.registration import register, registry
    except Exception:
        return False
    for env_id in ENV_IDS:
        if env_id not in registry:
            register(id=env_id, entry_point=lambda env_id=env_id, **kw: RobotTaskEnv(env_id, **kw))
    return True
