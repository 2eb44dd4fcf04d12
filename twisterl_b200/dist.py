"""Multi-GPU plumbing: one process per GPU, independent env shards, replicated weights.

The rollout itself has no exchange step (episodes are independent); per iteration there is one
broadcast of the flat fp32 weight blob from the trainer rank and one all-reduce of a small statistics vector.  On the
GPUs both are NCCL calls made by the library itself (`twr_broadcast_weights` / `twr_allreduce_stats` of the C ABI, on the
engine's stream) -- what a Rust or C++ host would call; `torch.distributed` (gloo, CPU) only carries the 128-byte NCCL
unique id from rank 0 to the other ranks, and stands in for NCCL in the CPU tests of this host logic.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

STATS_FIELDS = ("episodes", "successes", "reward_sum", "records")


def env_id_base(rank: int, num_episodes: int) -> int:
    """First global env id of this rank's shard: rank r owns [r*E, (r+1)*E).  The Philox streams are keyed
    by global env id, so a rollout does not depend on how many GPUs it was sharded over."""
    return int(rank) * int(num_episodes)


def _stats_dict(v) -> dict:
    return {"episodes": int(v[0]), "successes": int(v[1]), "reward_sum": float(v[2]), "records": int(v[3]),
            "success_rate": v[1] / max(v[0], 1.0), "mean_reward": v[2] / max(v[0], 1.0)}


class Comm:
    """The exchanges of one rank.  `engine` given: NCCL through the C ABI (the communicator lives in the engine);
    `engine=None`: torch.distributed's default group (gloo in the CPU tests).  world == 1: every call is a no-op."""

    def __init__(self, engine=None, rank: int | None = None, world: int | None = None):
        self.engine = engine
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else int(world)
        self.backend = "none"
        if self.world == 1:
            return
        import torch.distributed as dist
        if not dist.is_initialized():                  # bootstrap channel only (MASTER_ADDR / MASTER_PORT from the launcher)
            dist.init_process_group("gloo", rank=self.rank, world_size=self.world)
        self._dist = dist
        if engine is None:
            self.backend = "gloo"
            return
        from . import _lib
        if (engine.rank, engine.world) != (self.rank, self.world):
            raise ValueError("engine rank/world differ from the launcher's")
        L = _lib.load()
        uid = np.zeros(128, np.uint8)
        if self.rank == 0:
            _lib.check(L.twr_comm_unique_id(_lib.ptr(uid)))
        box = [uid.tobytes()]
        dist.broadcast_object_list(box, src=0)
        uid = np.frombuffer(box[0], np.uint8).copy()
        _lib.check(L.twr_comm_init(engine._h, _lib.ptr(uid)))
        self.backend = f"nccl {L.twr_comm_version()} via the C ABI"

    # ---- per-iteration exchanges -----------------------------------------------------------------
    def broadcast_weights(self, target, root: int = 0):
        """NCCL path: `target` is a policy handle (twr_policy*) -- rank `root`'s blob reaches every engine and the operand
        layouts are rebuilt.  gloo path: `target` is a torch tensor, broadcast in place."""
        if self.engine is not None:
            from . import _lib
            _lib.check(_lib.load().twr_broadcast_weights(self.engine._h, target, int(root)))
        elif self.world > 1:
            self._dist.broadcast(target, src=root)
        return target

    def allreduce(self, values, op: str = "sum") -> list:
        v = np.ascontiguousarray(values, dtype=np.float64)
        if self.world == 1:
            return v.tolist()
        if self.engine is not None:
            from . import _lib
            _lib.check(_lib.load().twr_allreduce_stats(self.engine._h, _lib.ptr(v), int(v.size), 1 if op == "max" else 0))
            return v.tolist()
        import torch
        t = torch.from_numpy(v)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX if op == "max" else self._dist.ReduceOp.SUM)
        return t.tolist()

    def allreduce_stats(self, episodes: int, successes: int, reward_sum: float, records: int) -> dict:
        return _stats_dict(self.allreduce([episodes, successes, reward_sum, records]))

    def max_over_ranks(self, x: float) -> float:
        return float(self.allreduce([x], op="max")[0])

    def barrier(self):
        if self.world > 1:
            self.allreduce([0.0])

    def close(self):
        if self.engine is not None and self.world > 1 and getattr(self.engine, "_h", None):
            from . import _lib
            _lib.load().twr_comm_destroy(self.engine._h)
        if self.world > 1 and self._dist.is_initialized():
            self._dist.destroy_process_group()
