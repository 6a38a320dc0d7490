"""CPU suite, part 4: the N>1 host logic on two gloo processes (no GPU): cloud sharding covers the
batch exactly once, the weak-scaling seeds are world-size independent, the timing reduction is a MAX
over ranks, and the training-time gradient all-reduce averages a flat bucket."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gridgcn_b200 import shard


def test_shard_range_partitions():
    for total in (0, 1, 7, 12, 96, 97):
        for world in (1, 2, 3, 8):
            blocks = [shard.shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_weak_scaling_seeds_are_world_independent():
    one = shard.cloud_seeds(8, 0) + shard.cloud_seeds(8, 1)
    assert one == list(range(16))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, w, _ = shard.env_rank_world()
        assert (r, w) == (rank, world)
        # sharded "clouds processed" sum to the batch; device time is the max over ranks
        a, b = shard.shard_range(13, rank, world)
        n = torch.tensor([b - a], dtype=torch.int64)
        dist.all_reduce(n)
        t = shard.max_over_ranks([1.0 + rank, 5.0 - rank])
        # gradient bucket all-reduce (mean)
        p1 = torch.nn.Parameter(torch.zeros(3))
        p2 = torch.nn.Parameter(torch.zeros(2, 2))
        p1.grad = torch.full((3,), float(rank + 1))
        p2.grad = torch.full((2, 2), float(10 * (rank + 1)))
        shard.allreduce_gradients([p1, p2])
        out[rank] = (int(n.item()), t, p1.grad.tolist(), p2.grad.flatten().tolist())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    for rank in range(world):
        n, t, g1, g2 = res[rank]
        assert n == 13
        assert t == [2.0, 5.0]
        assert np.allclose(g1, [1.5] * 3) and np.allclose(g2, [15.0] * 4)
