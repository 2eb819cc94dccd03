"""Device-resident rollout: policy -> env step -> replay write, without leaving the GPU.

The reference's training loop (train.py:39-60) is SB3's SAC.learn: `collect_rollouts` asks the policy for an action, steps
a DummyVecEnv of ONE env, and `ReplayBuffer.add` stores (obs, next_obs, action, reward, done) with next_obs replaced by
info["terminal_observation"] for finished envs and TimeLimit.truncated kept apart.  Here the same three stages run for
N envs per step as three stream-ordered pieces -- the caller's policy (any torch callable on the observation dict), the
step kernels, `urgym_replay_write` -- optionally captured into one CUDA graph.  SAC's gradient step is not part of this
repository (the caller); the ring is laid out so that a torch SAC can sample it by index."""
import ctypes
from typing import Callable, Dict, Optional

import torch

from . import _native as nat
from .vec_env import UR5VecEnv, _ptr


class DeviceReplayRing:
    """(obs, action, reward, next_obs, done, timeout) rings on the device; `capacity` rows, oldest overwritten."""

    def __init__(self, env: UR5VecEnv, capacity: int):
        if capacity < env.num_envs:
            raise ValueError("capacity must be at least num_envs")
        kw = dict(device=env.device)
        self.capacity, self.obs_dim = int(capacity), env.obs_dim
        self.obs = torch.zeros((capacity, env.obs_dim), dtype=torch.float32, **kw)
        self.next_obs = torch.zeros((capacity, env.obs_dim), dtype=torch.float32, **kw)
        self.actions = torch.zeros((capacity, 6), dtype=torch.float32, **kw)
        self.reward = torch.zeros(capacity, dtype=torch.float32, **kw)
        self.done = torch.zeros(capacity, dtype=torch.uint8, **kw)
        self.timeout = torch.zeros(capacity, dtype=torch.uint8, **kw)
        self.cursor = torch.zeros(1, dtype=torch.int64, **kw)        # transitions written so far (device-resident)

    def __len__(self) -> int:
        return min(int(self.cursor.item()), self.capacity)

    def sample(self, batch_size: int, generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
        """uniform minibatch (SB3 ReplayBuffer.sample): dones already exclude time-outs, as SAC's target needs"""
        idx = torch.randint(0, len(self), (batch_size,), device=self.obs.device, generator=generator)
        return {"observations": self.obs[idx], "next_observations": self.next_obs[idx], "actions": self.actions[idx],
                "rewards": self.reward[idx], "dones": (self.done[idx] & (1 - self.timeout[idx])).float()}


class Rollout:
    """env.step driven by `policy(obs_dict) -> actions [N,6]`, every transition appended to `ring`.

        ro = Rollout(env, policy, ring); ro.run(64)              # eager
        ro.capture(8); ro.replay(8)                              # 8 steps per CUDA-graph replay (policy must be capturable)
    """

    def __init__(self, env: UR5VecEnv, policy: Callable[[Dict[str, torch.Tensor]], torch.Tensor], ring: DeviceReplayRing):
        self.env, self.policy, self.ring = env, policy, ring
        self.prev = torch.empty_like(env.obs)           # the observation the action was computed from
        self.actions = torch.zeros((env.num_envs, 6), dtype=torch.float32, device=env.device)
        self.graph, self.graph_steps = None, 0

    def _one(self):
        env, r = self.env, self.ring
        self.prev.copy_(env.obs)
        self.actions.copy_(self.policy(env._obs))
        env.step(self.actions)
        nat.check(env.h, env.L.urgym_replay_write(
            env.h, _ptr(self.prev), _ptr(self.actions), _ptr(env.reward), _ptr(env.terminated), _ptr(env.truncated),
            _ptr(env.obs), _ptr(env.terminal_obs), _ptr(r.obs), _ptr(r.next_obs), _ptr(r.actions), _ptr(r.reward),
            _ptr(r.done), _ptr(r.timeout), ctypes.c_int64(r.capacity), _ptr(r.cursor), env._stream()))

    def run(self, steps: int) -> None:
        self.env._need_reset()
        for _ in range(steps):
            self._one()

    def capture(self, steps: int) -> None:
        self.env._need_reset()
        self._one()                                     # warm-up outside the capture (lazy initialisations of torch)
        torch.cuda.synchronize(self.env.device)
        self.graph, self.graph_steps = torch.cuda.CUDAGraph(), steps
        with torch.cuda.graph(self.graph):
            for _ in range(steps):
                self._one()

    def replay(self, times: int = 1) -> None:
        for _ in range(times):
            self.graph.replay()


def mlp_policy(weights: Dict[str, torch.Tensor], dtype: torch.dtype = torch.float32) -> Callable:
    """SB3 SAC MultiInputPolicy actor, deterministic: features = concat(achieved_goal, desired_goal, observation)
    (CombinedExtractor, keys sorted) -> 256 -> ReLU -> 256 -> ReLU -> mu -> tanh.  `weights`: the arrays of
    tests/golden/policy_*.npz as device tensors.  Plain torch matmuls (library GEMMs): the policy is the caller's."""
    w = {k: v.to(dtype) for k, v in weights.items()}

    def act(obs: Dict[str, torch.Tensor]) -> torch.Tensor:
        x = torch.cat([obs["achieved_goal"], obs["desired_goal"], obs["observation"]], dim=1).to(dtype)
        h = torch.relu(torch.addmm(w["latent_pi_0_bias"], x, w["latent_pi_0_weight"].T))
        h = torch.relu(torch.addmm(w["latent_pi_2_bias"], h, w["latent_pi_2_weight"].T))
        return torch.tanh(torch.addmm(w["mu_bias"], h, w["mu_weight"].T)).float()
    return act
