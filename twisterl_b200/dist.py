"""Multi-GPU plumbing: one process per GPU, independent env shards, replicated weights.

The rollout itself has no exchange step (episodes are independent); the only
collectives per iteration are a broadcast of the flat fp32 weight blob from the trainer rank and an
all-reduce of a four-number stats vector.  Both go through torch.distributed (NCCL on the GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

STATS_FIELDS = ("episodes", "successes", "reward_sum", "records")


def env_id_base(rank: int, num_episodes: int) -> int:
    """First global env id of this rank's shard: rank r owns [r*E, (r+1)*E).  The Philox streams are keyed
    by global env id, so a rollout does not depend on how many GPUs it was sharded over."""
    return int(rank) * int(num_episodes)


def broadcast_weights(blob: torch.Tensor, src: int = 0) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src=src)
    return blob


def allreduce_stats(episodes: int, successes: int, reward_sum: float, records: int, device=None) -> dict:
    t = torch.tensor([episodes, successes, reward_sum, records], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    v = t.tolist()
    return {"episodes": int(v[0]), "successes": int(v[1]), "reward_sum": float(v[2]), "records": int(v[3]),
            "success_rate": v[1] / max(v[0], 1.0), "mean_reward": v[2] / max(v[0], 1.0)}


def max_over_ranks(x: float, device=None) -> float:
    t = torch.tensor([x], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
