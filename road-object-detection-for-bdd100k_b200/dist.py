"""Image-sharded data parallelism for the box-level path.

Every image is independent in ARM / ODM target generation and in decode + NMS, so a batch is
partitioned into contiguous slices, one per rank (one process per GPU).  No input is exchanged
and the anchor table is regenerated locally.  The only collective is an all-gather of the
per-rank detection counts (`[C, B_local]` int32, what `streaming_tp_fp_arrays` needs to size its
global arrays, utils/tf_extended/metrics.py:178-194): NCCL over NVLink on GPUs, gloo in CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_images: int, rank: int, world: int):
    """Contiguous [begin, end) of the images rank `rank` owns; sizes differ by at most one."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world: %d/%d" % (rank, world))
    base, extra = divmod(n_images, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allgather_counts(counts: torch.Tensor, n_images: int, group=None) -> torch.Tensor:
    """counts: this rank's `[C, B_local]` int32 detection counts (B_local = its shard_range size).
    Returns the global `[C, n_images]` tensor on every rank.  Uneven shards are padded to the
    largest shard for the collective and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized():
        return counts
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_images, r, world) for r in range(world)]
    width = max(e - b for b, e in sizes)
    c = counts.shape[0]
    mine = counts.new_zeros((c, width))
    mine[:, :counts.shape[1]] = counts
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine.contiguous(), group=group)
    return torch.cat([p[:, :e - b] for p, (b, e) in zip(parts, sizes)], dim=1)
