"""Multi-GPU layout: one process per GPU, contiguous env index ranges per rank, no traffic in step()/reset().

The only collective is the sum of the eight per-shard episode statistics (episodes, return sum, length sum, successes,
collisions, truncations, env steps, reset iterations) -- 64 bytes, issued by the caller every K steps.  The reset
stream is keyed by (seed, GLOBAL env index, reset event), so an env's episodes do not depend on how many ranks the
job uses."""
from typing import Dict, Tuple

import torch
import torch.distributed as dist

from . import _native as nat


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of the envs rank `rank` owns; sizes differ by at most one"""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_stats(stats: Dict[str, float], device=None) -> Dict[str, float]:
    """sum the per-shard statistics over all ranks (NCCL on GPUs, gloo on CPU); identity without a process group"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(stats)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([stats[k] for k in nat.STAT_NAMES], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return dict(zip(nat.STAT_NAMES, t.tolist()))


def summarize(stats: Dict[str, float]) -> Dict[str, float]:
    ep = max(stats["episodes"], 1.0)
    return {"episodes": stats["episodes"], "mean_return": stats["return_sum"] / ep, "mean_length": stats["length_sum"] / ep,
            "success_rate": stats["successes"] / ep, "collision_rate": stats["collisions"] / ep,
            "truncation_rate": stats["truncations"] / ep, "env_steps": stats["env_steps"],
            "reset_iterations_per_episode": stats["reset_iterations"] / ep}
