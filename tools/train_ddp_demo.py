#!/usr/bin/env python
"""Data-parallel training demo (one process per GPU, NCCL): clouds sharded across ranks, CUDA index operators,
training-mode GridConv block, ONE flat gradient all-reduce per step (train.train_step).  Prints the mean loss
per step on rank 0 and checks that every rank holds identical parameters afterwards.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_ddp_demo.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridgcn_b200 as gg  # noqa: E402,F401
from gridgcn_b200 import shard, stack, synth, train  # noqa: E402

rank, world, local = shard.env_rank_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
per_rank = 8
cfg = stack.cls1024_4layer(16)
torch.manual_seed(0)  # same initial parameters on every rank
model = train.GridGcnClassifier(cfg, stack.init_params(cfg, seed=0), num_classes=4).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=2e-3)
seeds = shard.cloud_seeds(per_rank, rank)
data, npts = synth.make_batch(per_rank, cfg.num_points, seed0=seeds[0], voxels=cfg.voxels)
d, n = torch.from_numpy(data).to(dev), torch.from_numpy(npts).to(dev)
labels = torch.tensor([s % 4 for s in seeds], device=dev)
for it in range(steps):
    loss = train.train_step(model, opt, d, n, labels)
    mean = torch.tensor([loss], device=dev)
    if world > 1:
        dist.all_reduce(mean)
    if rank == 0 and (it % 10 == 0 or it == steps - 1):
        print("step %3d  mean loss over %d ranks %.4f" % (it, world, mean.item() / world), flush=True)
chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
if world > 1:
    dist.all_gather(allc, chk)
    assert all(torch.equal(allc[0], c) for c in allc), "parameters diverged across ranks"
if rank == 0:
    print("parameters identical on %d rank(s): checksum %.9f" % (world, chk.item()))
if world > 1:
    dist.destroy_process_group()
