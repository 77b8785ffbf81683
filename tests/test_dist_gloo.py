"""CPU test of the N > 1 host logic with the gloo backend, world_size 2."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    import vloam_b200  # noqa: F401
    from vloam_b200 import dist as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, lr = D.env_rank_world()
    mine = list(D.shard_streams(11, w, r))
    # this rank "processes" its streams for a rank-dependent time
    ms = 10.0 * (rank + 1)
    value = D.aggregate_throughput(len(mine) * 4, ms, dist)          # 4 steps
    mx = D.max_over_ranks(ms, dist)
    # point-sharded exchange: each rank holds partial normal equations of 3 streams
    part = torch.full((3, 28), float(rank + 1), dtype=torch.float64)
    D.allreduce_normal_equations(part, dist)
    # handle exchange of the point-sharded mode: 64 opaque bytes per rank, gathered in rank order
    handles = D.gather_ipc_handles(bytes([rank * 16 + (i % 16) for i in range(64)]), dist)
    assert len(handles) == 64 * world
    for rr in range(world):
        assert handles[64 * rr: 64 * rr + 64] == bytes([rr * 16 + (i % 16) for i in range(64)])
    dist.barrier()
    q.put((rank, mine, value, mx, part.sum().item()))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, v0, m0, e0), (r1, s1, v1, m1, e1) = out
    assert sorted(s0 + s1) == list(range(11)) and abs(len(s0) - len(s1)) <= 1      # a partition, balanced
    assert m0 == m1 == 20.0                                                        # max over ranks
    assert v0 == v1 == pytest.approx(11 * 4 / 0.020)                               # units of all ranks / slowest rank
    assert e0 == e1 == 3 * 28 * 3.0                                                # sum of the partial equations


def test_shard_streams_properties():
    import vloam_b200  # noqa: F401
    from vloam_b200 import dist as D
    for total in (0, 1, 7, 8, 129):
        for world in (1, 2, 3, 8):
            parts = [list(D.shard_streams(total, world, r)) for r in range(world)]
            assert sum(parts, []) == list(range(total))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        D.shard_streams(4, 2, 2)
