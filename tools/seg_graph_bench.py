"""Throughput of the full ScanNet-8192 graph (encoder + BallKNN decoder + head, the reference's shipped
config, BASELINE.json configs[2]) -- an extra data point for DESIGN.md, not the bench.py headline."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gridgcn_b200 import stack, synth

dev = torch.device("cuda:0")
cfg, up = stack.seg8192_shipped(), stack.UpCfg()
params = stack.init_seg_params(cfg, up, seed=0)
net = stack.GridGcnSeg(cfg, up, params, dev, precision="tf32x3")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for B in [int(a) for a in sys.argv[1:]] or [12, 96]:
    base, _ = synth.make_batch(min(B, 12), cfg.num_points, 0, voxels=cfg.voxels)
    data = torch.from_numpy(np.tile(base, ((B + len(base) - 1) // len(base), 1, 1))[:B].copy()).to(dev)
    npts = torch.full((B, 1), cfg.num_points, dtype=torch.int32, device=dev)
    for _ in range(3):
        net(data, npts)
    torch.cuda.synchronize()
    evs = []
    for _ in range(10):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); net(data, npts); e.record(); evs.append((s, e))
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in evs) / len(evs)
    print("seg8192_shipped full graph  B=%d  %.3f ms/step  %.3e points/s" % (B, ms, B * cfg.num_points / (ms * 1e-3)))
