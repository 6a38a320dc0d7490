"""Probe: how does tcgen05 kind::tf32 treat the low 13 mantissa bits of fp32 operands?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gridgcn_b200 as gg
dev = torch.device("cuda:0")
L = gg._lib.lib()
rng = np.random.default_rng(0)
N, K = 64, 8
A = rng.normal(size=(128, K)).astype(np.float32)
B = rng.normal(size=(N, K)).astype(np.float32)
a, b = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
d = torch.zeros((128, N), dtype=torch.float32, device=dev)
rc = L.gridgcn_debug_tc_gemm(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, 0, None)
torch.cuda.synchronize()
got = d.cpu().numpy().astype(np.float64)
def trunc(x): return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
def rna(x):
    u = x.view(np.uint32).astype(np.uint64) + 0x1000
    return (u.astype(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
for name, f in (("truncate", trunc), ("round-nearest-away", rna), ("exact fp32", lambda x: x)):
    want = f(A).astype(np.float64) @ f(B).astype(np.float64).T
    print("%-20s max |diff| = %.3e" % (name, np.abs(got - want).max()))
