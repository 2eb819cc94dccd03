"""Adapter that lets the shared parity harness drive the CUDA library through its public Python/C-ABI surface."""
import numpy as np
import torch

import urgym_b200 as ug


class GpuSim:
    def __init__(self, env_id, geom, n, seed=0, offset=0, autoreset=True, device=0, link_dist_mode=0):
        self.vec = ug.UR5VecEnv(env_id, n, device=device, seed=seed, env_index_offset=offset,
                                geometry="capsule" if geom == 1 else "hull", auto_reset=autoreset,
                                link_dist="workbench" if link_dist_mode else "obstacle")
        self.n = n

    def reset(self):
        return self.vec.reset()["observation"].cpu().numpy().copy()

    def step(self, actions):
        obs, rew, term, trunc, info = self.vec.step(torch.as_tensor(np.ascontiguousarray(actions, np.float32)).cuda(self.vec.device))
        torch.cuda.synchronize()
        return dict(obs=obs["observation"].cpu().numpy().copy(), reward=rew.cpu().numpy().copy(),
                    terminated=term.cpu().numpy().astype(bool), truncated=trunc.cpu().numpy().astype(bool),
                    is_success=info["is_success"].cpu().numpy().astype(bool),
                    terminal_obs=info["terminal_observation"].cpu().numpy().copy())

    # scenario injection (API layout), same names as tests/_hostcheck.HostCheckSim
    def _set(self, name, v):
        self.vec.set_state(name, torch.as_tensor(np.ascontiguousarray(v, np.float32)))

    def set_goal(self, g):
        self._set("goal", g)

    def set_obstacle(self, o):
        self._set("obstacle", o)

    def set_obstacle_end(self, o):
        self._set("obstacle_end", o)

    def set_obstacle_start(self, o):
        self._set("obstacle_start", o)

    def refresh(self):
        return self.vec.refresh().cpu().numpy().astype(bool)


class GpuSampleSim:
    """A full-size simulator seen through a sample of its envs: the harness drives `len(indices)` rows, every other env
    of the batch steps with its own random actions (generated on the device)."""

    def __init__(self, env_id, n_total, indices, seed=0, offset=0, device=0, geometry="capsule"):
        self.vec = ug.UR5VecEnv(env_id, n_total, device=device, seed=seed, env_index_offset=offset, geometry=geometry)
        self.idx = torch.as_tensor(np.asarray(indices, np.int64) - offset, device=self.vec.device)
        self.gen = torch.Generator(device=self.vec.device).manual_seed(1234)
        self.n_total = n_total

    def _rows(self, t):
        return t[self.idx].cpu().numpy().copy()

    def reset(self):
        return self._rows(self.vec.reset()["observation"])

    def step(self, actions):
        a = torch.rand((self.n_total, 6), device=self.vec.device, generator=self.gen) * 2.4 - 1.2
        a[self.idx] = torch.as_tensor(np.ascontiguousarray(actions, np.float32)).to(self.vec.device)
        obs, rew, term, trunc, info = self.vec.step(a)
        return dict(obs=self._rows(obs["observation"]), reward=self._rows(rew), terminated=self._rows(term).astype(bool),
                    truncated=self._rows(trunc).astype(bool), is_success=self._rows(info["is_success"]).astype(bool),
                    terminal_obs=self._rows(info["terminal_observation"]))
