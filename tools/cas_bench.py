"""Times Gridify_occaware (build + coverage-aware sampling + query) against Gridify at the seg-8192
first-layer shape.  Usage: python tools/cas_bench.py [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridgcn_b200 as gg  # noqa: E402
from gridgcn_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 192
dev = torch.device("cuda:0")
base, _ = synth.make_batch(min(B, 16), 8192, seed0=0, voxels=[0.05])
data = torch.from_numpy(np.tile(base, ((B + len(base) - 1) // len(base), 1, 1))[:B].copy()).to(dev)
npts = torch.full((B, 1), 8192, dtype=torch.int32, device=dev)
kw = dict(max_p_grid=64, max_o_grid=1024, kernel_size=3, loc=1, coord_shift=[1.0] * 3,
          voxel_size=[0.05] * 3, grid_size=[40] * 3)


def timed(f, reps=10):
    for _ in range(3):
        f()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        f()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


print("B=%d  Gridify %.3f ms   GridifyKNN %.3f ms   Gridify_occaware %.3f ms   occaware+knn %.3f ms" % (
    B, timed(lambda: gg.Gridify(data, npts, **kw)), timed(lambda: gg.GridifyKNN(data, npts, **kw)),
    timed(lambda: gg.Gridify_occaware(data, npts, seed=1, **kw)),
    timed(lambda: gg.Gridify_occaware(data, npts, seed=1, knn_query=True, **kw))))
