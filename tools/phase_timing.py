"""Debug helper: per-phase cycle breakdown of the per-edge tensor-core kernel (CTA 0), per layer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gridgcn_b200 as gg
from gridgcn_b200 import stack, synth
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
cfg = stack.seg8192_4layer(64)
params = stack.init_params(cfg, 0)
data, npts = synth.make_batch(min(B, 16), cfg.num_points, 0, voxels=cfg.voxels)
data = np.tile(data, ((B + 15) // 16, 1, 1))[:B]
enc = stack.GridGcnEncoder(cfg, params, dev, precision="tf32x3")
d, n = torch.from_numpy(data).to(dev), torch.full((B, 1), cfg.num_points, dtype=torch.int32, device=dev)
enc(d, n, keep_trace=True)
tr = enc.trace
L = gg._lib.lib()
buf = torch.zeros(8, dtype=torch.int64, device=dev)
names = ["gather", "sync0", "hidMMA", "hidEpi", "sync1", "finMMA", "finEpi", "sync2"]
for i, conv in enumerate(enc.convs):
    tin = d if i == 0 else tr[i - 1]["table"]
    buf.zero_()
    L.gridgcn_debug_phase_buffer(buf.data_ptr())
    conv(tin, tr[i]["nebidx"], tr[i]["cent"], tr[i]["centmsk"])
    torch.cuda.synchronize()
    L.gridgcn_debug_phase_buffer(None)
    v = buf.cpu().numpy().astype(np.float64)
    O, K = cfg.layers[i].max_o_grid, cfg.layers[i].max_p_grid
    print("layer %d: total %.0f kcycles; " % (i, v.sum() / 1e3) + "  ".join("%s %.1f%%" % (nm, 100 * x / max(v.sum(), 1)) for nm, x in zip(names, v)))
