#!/usr/bin/env python
"""Builds an experimental variant of libgridgcn_b200.so for A/B timing on the GPU box:

    python tools/build_variant.py NAME -DGG_KNN_MIN_CTAS=4 ...   ->  grid-gcn_b200/libgridgcn_b200_NAME.so
    GRIDGCN_B200_LIB=grid-gcn_b200/libgridgcn_b200_NAME.so python bench.py ...

Every .cu is recompiled with the extra flags into grid-gcn_b200/build_NAME/ (the product library is
not touched)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gridgcn_b200  # noqa: E402,F401
B = sys.modules["gridgcn_b200.build"]  # the package attribute `build` is the function

name, extra = sys.argv[1], sys.argv[2:]
objdir = os.path.join(B._HERE, "build_" + name)
os.makedirs(objdir, exist_ok=True)
procs, objs = [], []
for src in B.sources():
    obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
    objs.append(obj)
    procs.append(subprocess.Popen([B._nvcc()] + B.NVCC_FLAGS + extra + ["-c", src, "-o", obj]))
for p in procs:
    if p.wait() != 0:
        raise SystemExit("nvcc failed")
out = os.path.join(B._HERE, "libgridgcn_b200_%s.so" % name)
subprocess.check_call([B._nvcc(), "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
print(out)
