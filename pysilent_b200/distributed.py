"""Frame-sharded data parallelism: one process per GPU, frames (or camera streams) split across ranks, and ONE
collective per batch -- the variable-length gather of feature points.

The reference is single-process / single-GPU (``recognition_testing.py:64``); nothing in S1-S8 couples frames, so the
data path needs no exchange. Points are tiny (32 B each, a handful per level): the gather is latency-bound, so it is
issued as one padded ``all_gather`` per batch rather than per frame. Works on any ``torch.distributed`` backend
(``nccl`` on the GPU box over NVLink, ``gloo`` in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world_size):
    """Contiguous block of ``total`` frames owned by ``rank`` (block sizes differ by at most one)."""
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_points_padded(points, count, frame_offset, levels_per_frame, capacity, group=None):
    """Sync-free form for steady-state loops: every rank contributes exactly ``capacity`` rows.

    :param count: 1-element int64 DEVICE tensor (as written by ``silent_pipeline_run``).
    :return: ``(everyone [world, capacity, 4], counts [world])``; rows beyond ``counts[r]`` are padding.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    padded = points[:capacity].clone()
    padded[:, 0] += frame_offset * levels_per_frame
    counts = torch.empty(world, dtype=torch.int64, device=points.device)
    everyone = torch.empty((world, capacity, 4), dtype=torch.int64, device=points.device)
    if world == 1:
        counts.copy_(count.reshape(1))
        everyone[0].copy_(padded)
        return everyone, counts
    dist.all_gather_into_tensor(counts, count.reshape(1), group=group)
    dist.all_gather_into_tensor(everyone.view(world * capacity, 4), padded, group=group)
    return everyone, counts


def gather_points(points, count=None, frame_offset=0, levels_per_frame=1, capacity=None, group=None):
    """Gather every rank's feature points in global frame order.

    :param points: int64 ``[cap_or_K, 4]`` rows ``(local_level, y, x, 0)`` of this rank (row-major order).
    :param count: number of valid rows (int or 1-element tensor); default ``len(points)``.
    :param frame_offset: index of this rank's first frame in the global batch; level ids are rebased to
        ``global_frame * levels_per_frame + level`` so the concatenation is in the reference's row-major order
        (ranks hold contiguous frame blocks, see :func:`shard_range`).
    :param capacity: fixed padded row count per rank (same on every rank); default = max count over ranks
        (costs one extra tiny all-reduce).
    :return: ``(points [K_total, 4], counts [world_size])`` on every rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = points.device
    n = int(count) if count is not None else int(points.shape[0])
    n = min(n, int(points.shape[0]))
    local = points[:n].clone()
    local[:, 0] += frame_offset * levels_per_frame
    if world == 1:
        return local, torch.tensor([n], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    cap = int(capacity) if capacity is not None else int(counts.max().item())
    cap = max(cap, 1)
    padded = torch.zeros((cap, 4), dtype=torch.int64, device=dev)
    padded[:min(n, cap)] = local[:cap]
    everyone = torch.empty((world * cap, 4), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(everyone, padded, group=group)
    host_counts = counts.tolist()
    parts = [everyone[r * cap: r * cap + min(int(host_counts[r]), cap)] for r in range(world)]
    return torch.cat(parts, dim=0), counts
