"""stable-baselines3 `VecEnv`-shaped adapter over UR5VecEnv: what train.py's `SAC("MultiInputPolicy", env, ...)` sees
through `DummyVecEnv([lambda: Monitor(gym.make(id))])` (train.py:39-60; SB3 vec_env/dummy_vec_env.py semantics):

    obs = venv.reset()                                   # dict of numpy arrays [n_envs, ...]
    obs, rewards, dones, infos = venv.step(actions)      # infos[i]: "is_success", "TimeLimit.truncated", and on done
                                                         # "terminal_observation" (dict) and Monitor's "episode" {r, l}

stable-baselines3 itself is not a dependency (it is absent from the build image): the class implements the VecEnv
protocol (num_envs, observation_space, action_space, reset, step_async, step_wait, step, close, env_is_wrapped,
get_attr, set_attr, env_method, seed) without subclassing it.  Per-env Python dicts are built only for finished envs,
so the adapter stays usable at tens of thousands of envs; device-resident training loops should use UR5VecEnv directly."""
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from .envs import Box, DictSpace
from .vec_env import UR5VecEnv


class SB3VecEnvAdapter:
    def __init__(self, env_id: str, n_envs: int, device: int = 0, seed: int = 0, geometry: str = "capsule"):
        self.vec = UR5VecEnv(env_id, n_envs, device=device, seed=seed, geometry=geometry)
        self.num_envs = n_envs
        D, G = self.vec.obs_dim, self.vec.goal_dim
        self.observation_space = DictSpace(dict(observation=Box(-10.0, 10.0, shape=(D,), dtype=np.float32),
                                                desired_goal=Box(-10.0, 10.0, shape=(G,), dtype=np.float32),
                                                achieved_goal=Box(-10.0, 10.0, shape=(G,), dtype=np.float32)))
        self.action_space = Box(-1.0, 1.0, shape=(6,), dtype=np.float32)
        self.render_mode = None
        self._actions: Optional[np.ndarray] = None
        self._ep_ret = np.zeros(n_envs, np.float64)
        self._ep_len = np.zeros(n_envs, np.int64)
        self._host = self.vec.alloc_host_buffers(terminal_obs=True)

    # ---- VecEnv protocol
    def _obs_np(self, obs: np.ndarray) -> Dict[str, np.ndarray]:
        G = self.vec.goal_dim
        return {"observation": obs.copy(), "achieved_goal": obs[:, :G].copy(), "desired_goal": obs[:, 12:12 + G].copy()}

    def reset(self) -> Dict[str, np.ndarray]:
        obs = self.vec.reset()["observation"].cpu().numpy()
        self._ep_ret[:] = 0; self._ep_len[:] = 0
        return self._obs_np(obs)

    def step_async(self, actions: np.ndarray) -> None:
        self._actions = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.num_envs, 6)

    def step_wait(self):
        h = self._host
        h["actions"].copy_(torch.from_numpy(self._actions))
        self.vec.step_host(h["actions"], h)                     # host buffers in and out, one synchronisation
        obs, rew = h["obs"].numpy(), h["reward"].numpy().copy()
        term, trunc, succ = h["terminated"].numpy().astype(bool), h["truncated"].numpy().astype(bool), h["is_success"].numpy().astype(bool)
        dones = term | trunc
        self._ep_ret += rew; self._ep_len += 1
        G = self.vec.goal_dim
        infos: List[Dict[str, Any]] = [{"is_success": bool(succ[i]), "TimeLimit.truncated": bool(trunc[i] and not term[i])}
                                       for i in range(self.num_envs)]
        tobs = h["terminal_obs"].numpy()
        for i in np.nonzero(dones)[0]:
            row = tobs[i]
            infos[i]["terminal_observation"] = {"observation": row.copy(), "achieved_goal": row[:G].copy(),
                                                "desired_goal": row[12:12 + G].copy()}
            infos[i]["episode"] = {"r": float(self._ep_ret[i]), "l": int(self._ep_len[i])}     # Monitor (train.py:52)
            self._ep_ret[i] = 0.0; self._ep_len[i] = 0
        return self._obs_np(obs), rew, dones, infos

    def step(self, actions: np.ndarray):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        self.vec.close()

    def seed(self, seed: Optional[int] = None) -> Sequence[Optional[int]]:
        if seed is not None:
            self.vec.reseed(seed)
        return [seed] * self.num_envs

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        return [False] * self.num_envs

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self, attr_name, None)] * n

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        setattr(self, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> List[Any]:
        if method_name == "compute_reward":
            raise NotImplementedError("compute_reward needs the simulator's link distances; use UR5VecEnv rewards")
        raise AttributeError(method_name)
