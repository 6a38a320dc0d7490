"""Builds libgridgcn_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

The library is a plain C-ABI shared object (include/gridgcn_b200.h); it is git-ignored but travels
to the GPU box with the repository snapshot.
"""
import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libgridgcn_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    # NO --use_fast_math: voxelisation and distances need IEEE fp32 (SURVEY.md s8c rule 2, 10)
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libgridgcn_b200.so cannot be built")
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inc")) + \
        [os.path.join(_HERE, "..", "include", "gridgcn_b200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def _deps(path, seen=None):
    """The file and every quoted #include reachable from it (so that an edit recompiles only the
    objects that see it)."""
    import re
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path) as f:
        for inc in re.findall(r'^\s*#include\s+"([^"]+)"', f.read(), flags=re.M):
            _deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link libgridgcn_b200.so.  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and \
                all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in _deps(src)):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out.decode(errors="replace"))
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    # cudart is linked statically (nvcc default): the library has no load-time dependency besides
    # libstdc++/libc, so it loads on a CPU-only box for the symbol-export test.
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
