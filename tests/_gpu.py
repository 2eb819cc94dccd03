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
