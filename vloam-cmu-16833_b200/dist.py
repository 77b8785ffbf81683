"""Multi-GPU host logic: one process per GPU, streams sharded across ranks.

The path shards by *stream* — independent sensor sequences share nothing, so there is no data-path collective; the
only collectives are the bookkeeping ones below (barrier, max-over-ranks of the timed region, gather of per-rank
counters).  Backend: NCCL on the GPUs, gloo in the CPU tests (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import os

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_streams(total_streams: int, world: int, rank: int) -> range:
    """Global stream ids owned by `rank`: contiguous, balanced to within one stream, a partition of range(total)."""
    if not (0 <= rank < world) or total_streams < 0:
        raise ValueError((total_streams, world, rank))
    base, extra = divmod(total_streams, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def stream_seed(bench_seed: int, global_stream: int, n_base: int) -> int:
    """Seed of the synthetic base sequence a global stream replays (distinct sequences on distinct ranks)."""
    return bench_seed + (global_stream % n_base) + 16 * (global_stream // max(1, n_base) % 64)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Device-timed milliseconds -> max over ranks (the job is as slow as its slowest rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank: float, ms_this_rank: float, dist=None, device=None) -> float:
    """Whole-job units/s = (sum over ranks of the units processed) / (max over ranks of the timed region)."""
    total = sum_over_ranks(units_this_rank, dist, device)
    ms = max_over_ranks(ms_this_rank, dist, device)
    return total / (ms * 1e-3) if ms > 0 else 0.0


def allreduce_normal_equations(buf, dist=None):
    """Sum the per-rank partial normal equations [B, 28] = (21 upper J'J, 6 J'r, cost) in place.

    This is the one real exchange step the path has when the residuals of one stream are split across ranks
    (SURVEY.md section 8e layout (ii)); with stream sharding it is not needed.  Works on CPU (gloo) and GPU (NCCL)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return buf
    assert buf.shape[-1] == 28
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf


def gather_ipc_handles(local_handle: bytes, dist, device=None) -> bytes:
    """All-gather the 64-byte cudaIpcMemHandle_t of every rank's exchange buffer -> world * 64 bytes, rank order.
    Works over NCCL (device tensors) and gloo (CPU tensors, used by the CPU test)."""
    import torch
    assert len(local_handle) == 64
    t = torch.tensor(list(local_handle), dtype=torch.uint8, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)


def enable_point_sharding(lom, dist, device=None):
    """Switch a LidarOdometryMapping handle to the point-sharded solve (SURVEY.md section 8e layout (ii)): every rank must
    hold the same streams and feed the same scans; afterwards laserOdometryIO associates and accumulates only this rank's
    slice of the correspondences and the solve kernel sums the normal equations across ranks through peer memory."""
    rank, world = dist.get_rank(), dist.get_world_size()
    handles = gather_ipc_handles(lom.shard_ipc_handle(), dist, device)
    lom.shard_open_ipc(rank, world, handles)
    dist.barrier()          # nobody starts exchanging before every rank has mapped and cleared its slots


def enable_point_sharding_nccl(lom, dist, device=None):
    """The same split with the exchange done by ncclAllReduce between a wide accumulate launch and the step launch
    (BASELINE configs[4] as worded; include/vloam_b200.h vloam_shard_nccl_*).  Rank 0 creates the ncclUniqueId, everybody
    receives it through the job's process group, then the library builds its own communicator (collective)."""
    import torch
    from . import shard_nccl_unique_id
    rank, world = dist.get_rank(), dist.get_world_size()
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        t.copy_(torch.tensor(list(shard_nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    lom.shard_nccl_init(rank, world, bytes(t.cpu().numpy().tobytes()))
    dist.barrier()
