#!/usr/bin/env python
"""Data-parallel training demo (one process per GPU, NCCL): clouds sharded across ranks, CUDA index operators,
training-mode GridConv block, ONE flat gradient all-reduce per step (train.train_step).  Prints the mean loss
per step on rank 0 and checks that every rank holds identical parameters afterwards.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_ddp_demo.py [steps]
      [--block cuda]   forward + backward of the block on the library's own kernels (train_cuda.py) instead of torch ops
      [--graph]        the step replayed as CUDA graphs around the one eager all-reduce (train.GraphedTrainStep)
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
args = [a for a in sys.argv[1:] if not a.startswith("--")]
block = "cuda" if "--block" in sys.argv and sys.argv[sys.argv.index("--block") + 1] == "cuda" else "torch"
args = [a for a in args if a not in ("cuda", "torch")]
use_graph = "--graph" in sys.argv
steps = int(args[0]) if args else 40
per_rank = 8
cfg = stack.cls1024_4layer(16)
torch.manual_seed(0)  # same initial parameters on every rank
model = train.GridGcnClassifier(cfg, stack.init_params(cfg, seed=0), num_classes=4, block=block).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=2e-3, capturable=use_graph)
seeds = shard.cloud_seeds(per_rank, rank)
data, npts = synth.make_batch(per_rank, cfg.num_points, seed0=seeds[0], voxels=cfg.voxels)
d, n = torch.from_numpy(data).to(dev), torch.from_numpy(npts).to(dev)
labels = torch.tensor([s % 4 for s in seeds], device=dev)
gstep = train.GraphedTrainStep(model, opt, d, n, labels) if use_graph else None
if rank == 0:
    print("block = %s, %s" % (block, "CUDA-graph replay" if use_graph else "eager launches"), flush=True)
for it in range(steps):
    loss = float(gstep()) if use_graph else train.train_step(model, opt, d, n, labels)
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
