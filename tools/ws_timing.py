"""Debug helper for the warp-specialised first-layer kernel built with -DGG_WS_TIMING
(python tools/build_variant.py timing -DGG_WS_TIMING; GRIDGCN_B200_LIB=grid-gcn_b200/libgridgcn_b200_timing.so
python tools/ws_timing.py [B]): where does every role of CTA 0 wait?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gridgcn_b200 as gg
from gridgcn_b200 import stack, synth
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
cfg = stack.seg8192_4layer(64)
params = stack.init_params(cfg, 0)
data, npts = synth.make_batch(min(B, 16), cfg.num_points, 0, voxels=cfg.voxels)
data = np.tile(data, ((B + 15) // 16, 1, 1))[:B]
enc = stack.GridGcnEncoder(cfg, params, dev, precision="tf32x3")
d, n = torch.from_numpy(data).to(dev), torch.full((B, 1), cfg.num_points, dtype=torch.int32, device=dev)
enc(d, n, keep_trace=True)
tr = enc.trace[0]
L = gg._lib.lib()
buf = torch.zeros(32, dtype=torch.int64, device=dev)
for _ in range(2):
    buf.zero_()
    L.gridgcn_debug_phase_buffer(buf.data_ptr())
    enc.convs[0](d, tr["nebidx"], tr["cent"], tr["centmsk"])
    torch.cuda.synchronize()
    L.gridgcn_debug_phase_buffer(None)
v = buf.cpu().numpy().astype(np.float64)
units = max(v[31], 1)
waits = ["G:x0 slot free", "E0:img_free", "E0:s0_done", "EH:h_done", "EF:f_full", "EF:g_full", "MS:x0_full", "MS:acc_free",
         "MH:e0_done", "MA:eh_done", "MA:fg_free", "MF0:eh_done", "MF0:fg_free"]
roles = ["EF1", "E0", "EH", "EF0", "MS", "MH", "MA", "MF0", "MF1", "G"]
print("units of CTA 0: %d" % units)
for i, r in enumerate(roles):
    print("role %-3s total %8.0f cycles / unit" % (r, v[16 + i] / units))
for i, w in enumerate(waits):
    print("  wait %-16s %8.0f cycles / unit" % (w, v[i] / units))
