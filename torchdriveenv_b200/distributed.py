"""Env sharding across the GPUs of one box and the only collective on this path: the all-reduce of
the episode-statistics vector (what EvalNTimestepsCallback aggregates, examples/rl_training.py:99-108).
One process per GPU; nothing is exchanged on the step path."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ._capi import STAT_NAMES, TDE_NUM_STATS


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous env range [lo, hi) owned by `rank`; sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_episode_stats(local_stats, device=None) -> np.ndarray:
    """Sum the TDE_NUM_STATS vector over all ranks (NCCL on GPUs, gloo on CPU)."""
    t = torch.as_tensor(np.asarray(local_stats, np.float64).reshape(TDE_NUM_STATS).copy())
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def summarize(stats: np.ndarray) -> Dict[str, float]:
    s = {n: float(stats[i]) for i, n in enumerate(STAT_NAMES)}
    ep = max(s["episodes"], 1.0)
    return dict(episodes=s["episodes"], steps=s["steps"], mean_return=s["return_sum"] / ep, mean_length=s["length_sum"] / ep,
                offroad_rate=s["offroad"] / ep, collision_rate=s["collision"] / ep,
                traffic_light_violation_rate=s["traffic_light_violation"] / ep, success_rate=s["success"] / ep,
                mean_reached_waypoints=s["reached_waypoints"] / ep)
