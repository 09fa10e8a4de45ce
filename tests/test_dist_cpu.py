"""World-size-2 gloo test of the image sharding + count all-gather (host logic; no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rodet_b200.dist import allgather_counts, shard_range
    b, e = shard_range(n_images, rank, world)
    full = (torch.arange(11 * n_images, dtype=torch.int32).reshape(11, n_images) * 7) % 201
    got = allgather_counts(full[:, b:e].contiguous(), n_images)
    out[rank] = bool(torch.equal(got, full)) and got.shape == (11, n_images)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [8, 7])
def test_allgather_counts_world2(n_images):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n_images, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_shard_range_partitions():
    from rodet_b200.dist import shard_range
    for n in (0, 1, 7, 64, 255, 256):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker_tpfp(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rodet_b200.dist import allgather_tp_fp
    g = torch.Generator().manual_seed(5)
    n = 101
    scores = (torch.rand(n, generator=g) * 20).round() / 20           # many ties
    tp = torch.rand(n, generator=g) < 0.5
    fp = ~tp
    ids = torch.arange(n, dtype=torch.int64) * 3
    # interleaved ownership with uneven lengths: rank 0 holds 2/3 of the entries
    own = (torch.arange(n) % 3 != 2) if rank == 0 else (torch.arange(n) % 3 == 2)
    nobj = torch.tensor(7 if rank == 0 else 5, dtype=torch.int64)
    val = (nobj, torch.tensor(int(own.sum()), dtype=torch.int32), tp[own], fp[own], scores[own], ids[own])
    o, d, t, f, s = allgather_tp_fp(val)
    out[rank] = bool(int(o) == 12 and int(d) == n and torch.equal(t, tp) and torch.equal(f, fp) and torch.equal(s, scores))
    dist.destroy_process_group()


def test_allgather_tp_fp_world2():
    """Variable-length TP/FP/score arrays merged back into single-process order (f-2 exchange step)."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_tpfp, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] and out[1]
