"""Runs ONE forward of a BASELINE configuration after warm-up (for `ncu --metrics gpu__time_duration.sum`
launch lists): python tools/config_launches.py seg_graph|seg81920|cls1024 [B]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gridgcn_b200 import stack, synth

what = sys.argv[1] if len(sys.argv) > 1 else "seg_graph"
dev = torch.device("cuda:0")
if what == "seg_graph":
    cfg, up, B = stack.seg8192_shipped(), stack.UpCfg(), int(sys.argv[2]) if len(sys.argv) > 2 else 12
    net = stack.GridGcnSeg(cfg, up, stack.init_seg_params(cfg, up, seed=0), dev, precision="tf32x3")
elif what == "seg81920":
    cfg, B = stack.seg81920_shipped("gridify"), int(sys.argv[2]) if len(sys.argv) > 2 else 3
    net = stack.GridGcnEncoder(cfg, stack.init_params(cfg, seed=0), dev, precision="tf32x3")
else:
    cfg, B = stack.cls1024_4layer(32), int(sys.argv[2]) if len(sys.argv) > 2 else 32
    net = stack.GridGcnEncoder(cfg, stack.init_params(cfg, seed=0), dev, precision="tf32x3")
base, _ = synth.make_batch(min(B, 8), cfg.num_points, 0, voxels=cfg.voxels)
data = torch.from_numpy(np.tile(base, ((B + len(base) - 1) // len(base), 1, 1))[:B].copy()).to(dev)
npts = torch.full((B, 1), cfg.num_points, dtype=torch.int32, device=dev)
for _ in range(3):
    net(data, npts)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net(data, npts)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
