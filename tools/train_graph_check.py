#!/usr/bin/env python
"""config-4 training step alone (bench.measure_train_step): eager vs CUDA-graph replay, both blocks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench

world = int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
for block in ("cuda", "torch"):
    r = bench.measure_train_step(torch, dist, dev, world, block=block)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(r))
if world > 1:
    dist.destroy_process_group()
