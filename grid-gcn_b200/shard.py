"""Multi-GPU host logic: one process per GPU, clouds sharded across ranks, no data-path collective.

Every operator of the path is per-cloud (the reference indexes by i_batch with no cross-batch state,
gridify.cu:127,143; k_nn-inl.h:46; utils/ops.py:90), so a batch shards by cloud and the forward
path needs NO collective.  The only collectives are (i) the max-over-ranks of the device time when
benchmarking and (ii), for training, one all-reduce of the flattened gradient bucket per step
(the reference's analogue is MXNet's `kvstore: local`, segmentation/configs/configs.yaml:3).
A cloud is never split across devices (81 920 points are 1.3 MB).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(total, rank, world):
    """Contiguous block of clouds owned by `rank`: sizes differ by at most one, blocks are disjoint
    and cover [0, total)."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def cloud_seeds(clouds_per_rank, rank):
    """Weak scaling: rank r generates clouds with global ids [r*B, (r+1)*B) so that the union over
    ranks is the same data set whatever the world size."""
    return list(range(rank * clouds_per_rank, (rank + 1) * clouds_per_rank))


def max_over_ranks(values, device=None):
    """Element-wise MAX all-reduce of a list of python floats (device times of the timed region)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def allreduce_gradients(params):
    """Training-time collective: ONE all-reduce over a flat bucket of every gradient (mean)."""
    params = list(params)
    if not params or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    # the bucket layout must be the same on every rank: a parameter without a gradient on THIS rank (unused
    # branch, zero_grad(set_to_none=True)) contributes zeros instead of being skipped
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
