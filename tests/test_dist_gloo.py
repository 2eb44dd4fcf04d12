"""world_size-2 gloo test of the N>1 host path (env-id sharding, weight broadcast, stats reduction)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from twisterl_b200 import dist as twd
    comm = twd.Comm()                                        # no engine: torch.distributed (gloo) stands in for NCCL
    assert comm.backend == "gloo" and (comm.rank, comm.world) == (rank, world)
    blob = torch.full((1000,), float(rank + 1))
    comm.broadcast_weights(blob, root=0)
    stats = comm.allreduce_stats(episodes=10, successes=3 + rank, reward_sum=1.5 * (rank + 1), records=100 + rank)
    mx = comm.max_over_ranks(float(rank) + 0.25)
    comm.barrier()
    base = twd.env_id_base(rank, 65536)
    q.put((rank, float(blob.sum()), stats, mx, base))
    comm.close()


def test_two_rank_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, s, stats, mx, base in res:
        assert s == 1000.0                                   # rank 0's weights everywhere
        assert stats["episodes"] == 20 and stats["successes"] == 7 and stats["records"] == 201
        assert abs(stats["reward_sum"] - 4.5) < 1e-12 and abs(stats["success_rate"] - 0.35) < 1e-12
        assert mx == 1.25
        assert base == rank * 65536                          # disjoint Philox env-id ranges


def test_single_process_is_a_noop():
    from twisterl_b200 import dist as twd
    comm = twd.Comm(rank=0, world=1)
    b = torch.arange(4.0)
    assert comm.broadcast_weights(b) is b
    assert comm.allreduce_stats(2, 1, 0.5, 7)["records"] == 7
    assert comm.max_over_ranks(3.0) == 3.0
    comm.barrier(); comm.close()
