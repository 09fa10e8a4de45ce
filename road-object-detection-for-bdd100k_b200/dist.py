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


class PendingCounts:
    """Handle of an in-flight count all-gather; `result()` waits and returns the `[C, n_images]` tensor."""

    def __init__(self, work, parts, sizes, even):
        self._work, self._parts, self._sizes, self._even = work, parts, sizes, even

    def result(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._even:                                   # parts: [world, C, B_local]
            w, c, bl = self._parts.shape
            return self._parts.permute(1, 0, 2).reshape(c, w * bl)
        return torch.cat([p[:, :e - b] for p, (b, e) in zip(self._parts, self._sizes)], dim=1)


def allgather_counts(counts: torch.Tensor, n_images: int, group=None, async_op: bool = False):
    """counts: this rank's `[C, B_local]` int32 detection counts (B_local = its shard_range size).
    Returns the global `[C, n_images]` tensor on every rank (or, with `async_op=True`, a
    `PendingCounts` whose `result()` yields it, so the collective runs on NCCL's stream while the
    next batch is processed).  Even shards use one `all_gather_into_tensor`; uneven shards are
    padded to the largest shard and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized():
        return PendingCounts(None, counts.unsqueeze(0), None, True) if async_op else counts
    world = dist.get_world_size(group)
    sizes = [shard_range(n_images, r, world) for r in range(world)]
    width = max(e - b for b, e in sizes)
    c = counts.shape[0]
    even = (all(e - b == width for b, e in sizes) and counts.shape[1] == width
            and dist.get_backend(group) == "nccl")       # gloo: list form only
    if even:
        out = counts.new_empty((world, c, width))
        work = dist.all_gather_into_tensor(out, counts.contiguous(), group=group, async_op=True)
        pending = PendingCounts(work, out, sizes, True)
    else:
        mine = counts.new_zeros((c, width))
        mine[:, :counts.shape[1]] = counts
        parts = [torch.empty_like(mine) for _ in range(world)]
        work = dist.all_gather(parts, mine, group=group, async_op=True)
        pending = PendingCounts(work, parts, sizes, False)
    return pending if async_op else pending.result()


def allgather_tp_fp(value, group=None):
    """The exchange step of a sharded evaluation: merges the per-rank arrays accumulated by
    `tfe.streaming_tp_fp_arrays` (`StreamingTpFp.value(row, with_ids=True)` =
    (num_gbboxes, num_detections, tp, fp, scores, ids)) into the arrays a single process would hold.
    Lengths differ per rank: the counts are all-gathered first, the arrays are padded to the longest
    and all-gathered (NCCL over NVLink on GPUs), then put back into global detection order by `ids`
    (ties between equal scores are broken by position in `precision_recall`, so order matters).
    Returns (num_gbboxes, num_detections, tp, fp, scores) on every rank."""
    nobj, ndet, tp, fp, scores, ids = value
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        order = torch.argsort(ids, stable=True)
        return nobj, ndet, tp[order], fp[order], scores[order]
    world = dist.get_world_size(group)
    dev = scores.device
    n_local = torch.tensor([scores.numel(), int(nobj)], dtype=torch.int64, device=dev)
    sizes = [torch.empty_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    counts = [int(s[0]) for s in sizes]
    width = max(counts)
    # one padded [3, width] int64 block per rank: score bits, tp | fp << 1, ids
    mine = torch.zeros((3, width), dtype=torch.int64, device=dev)
    n = scores.numel()
    mine[0, :n] = scores.contiguous().view(torch.int32).to(torch.int64)
    mine[1, :n] = tp.to(torch.int64) | (fp.to(torch.int64) << 1)
    mine[2, :n] = ids
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    cat = torch.cat([p[:, :c] for p, c in zip(parts, counts)], dim=1)
    order = torch.argsort(cat[2], stable=True)
    cat = cat[:, order]
    total_obj = sum(int(s[1]) for s in sizes)
    return (torch.tensor(total_obj, dtype=torch.int64, device=dev), torch.tensor(cat.shape[1], dtype=torch.int32, device=dev),
            (cat[1] & 1).bool(), ((cat[1] >> 1) & 1).bool(), cat[0].to(torch.int32).view(torch.float32))
