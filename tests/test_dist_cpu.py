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
