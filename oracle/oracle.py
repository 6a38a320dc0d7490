"""ctypes front-end of the CPU oracle (oracle/gridgcn_oracle.c).

TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import this module.  Parity status: see the header of gridgcn_oracle.c
(the grid operators are pinned against the reference's own kernel bodies, oracle/_ref; coverage-aware sampling
and the numpy GridConv block are unpinned).

All functions take and return numpy arrays and mirror the reference operators' argument
names (gridifyop/gridify-inl.h:58-87,146-152; gridify_up-inl.h:58-81,137-143;
k_nn.cc:23-57; ball_k_nn-inl.h:33-39).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgridgcn_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile the C restatement with gcc (no GPU, no reference sources needed)."""
    src = os.path.join(_HERE, "gridgcn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libgridgcn_oracle.so"])
    return _LIB_PATH


def build_ref():
    """Compile oracle/_ref from the reference sources (only where /root/reference exists)."""
    if not os.path.isdir("/root/reference/gridifyop"):
        return None
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])  # libknn_ref.so and libgridify_ref.so
    return os.path.join(_HERE, "_ref", "libknn_ref.so")


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.gridgcn_oracle_set_threads.argtypes = [ctypes.c_int]
        _lib.gridgcn_oracle_max_threads.restype = ctypes.c_int
    return _lib


def set_k1_seconds(seconds):
    """Test-only: replay K1's time-seeded reservoirs literally with this `tv_usec` (None / negative: the
    canonical keep-first rule).  Used to compare with the reference's own kernel bodies in the overflow regime."""
    L = lib()
    L.gridgcn_oracle_set_k1_seconds.argtypes = [ctypes.c_longlong]
    L.gridgcn_oracle_set_k1_seconds(-1 if seconds is None else int(seconds))


def set_threads(n):
    lib().gridgcn_oracle_set_threads(int(n))


def max_threads():
    return int(lib().gridgcn_oracle_max_threads())


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def _triple(v, dtype):
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=dtype), (3,)))
    return a


def _gridify(fn, data, actual_numpoints, max_p_grid, max_o_grid, kernel_size, loc, coord_shift,
             voxel_size, grid_size, mode):
    data, pd = _f(data)
    assert data.ndim == 3 and data.shape[2] == 4, "data should be (B, N, 4)"
    B, N, _ = data.shape
    npts, pn = _i(np.asarray(actual_numpoints).reshape(B))
    O, P = int(max_o_grid), int(max_p_grid)
    shift = _triple(coord_shift, np.float32)
    voxel = _triple(voxel_size, np.float32)
    grid = _triple(grid_size, np.int32)
    nebidx = np.empty((B, O, P), np.int32)
    nebmsk = np.empty((B, O, P), np.float32)
    cent = np.empty((B, O, 4), np.float32)
    centmsk = np.empty((B, O), np.float32)
    centnum = np.empty((B, 1), np.int32)
    rc = fn(pd, pn, B, N, O, P, int(kernel_size), int(loc), shift.ctypes.data_as(_f32p),
            voxel.ctypes.data_as(_f32p), grid.ctypes.data_as(_i32p), int(mode),
            nebidx.ctypes.data_as(_i32p), nebmsk.ctypes.data_as(_f32p),
            cent.ctypes.data_as(_f32p), centmsk.ctypes.data_as(_f32p),
            centnum.ctypes.data_as(_i32p))
    if rc != 0:
        raise ValueError("oracle rejected the arguments (rc=%d)" % rc)
    return nebidx, nebmsk, cent, centmsk, centnum


def gridify(data, actual_numpoints, *, max_p_grid, max_o_grid, kernel_size, stride=1, loc=0,
            coord_shift=(0, 0, 0), voxel_size=(1, 1, 1), grid_size=(1, 1, 1),
            strict_reservoir=True):
    """Gridify (gridify-inl.h:99-128, gridify.cu:102-291).  ``stride`` is accepted and ignored,
    as in the reference (gridify.cu:112, never read).  ``strict_reservoir`` (default, like the product
    operator): K2's schedule-independent reservoir over the candidates beyond max_p_grid (gridify.cu:259-270);
    False = the canonical keep-first rule."""
    fn = lib().gridgcn_oracle_gridify
    fn.restype = ctypes.c_int
    return _gridify(fn, data, actual_numpoints, max_p_grid, max_o_grid, kernel_size, loc,
                    coord_shift, voxel_size, grid_size, 1 if strict_reservoir else 0)


def gridify_knn(data, actual_numpoints, *, max_p_grid, max_o_grid, kernel_size, stride=1, loc=0,
                coord_shift=(0, 0, 0), voxel_size=(1, 1, 1), grid_size=(1, 1, 1),
                dist_fma=False):
    """GridifyKNN (gridifyknn-inl.h, gridifyknn.cu:115-333)."""
    fn = lib().gridgcn_oracle_gridify_knn
    fn.restype = ctypes.c_int
    return _gridify(fn, data, actual_numpoints, max_p_grid, max_o_grid, kernel_size, loc,
                    coord_shift, voxel_size, grid_size, 1 if dist_fma else 0)


def gridify_occaware(data, actual_numpoints, *, max_p_grid, max_o_grid, kernel_size, stride=1, loc=0,
                     coord_shift=(0, 0, 0), voxel_size=(1, 1, 1), grid_size=(1, 1, 1), seed=0,
                     knn_query=False, dist_fma=False):
    """Gridify_occaware = Gridify with Coverage-Aware Sampling of the centre voxels (binary-only in
    the reference: additional.so, kernels gridify_kernel_build_index_occaware /
    gridify_occaware_sampling; restated from the sm_75 SASS, see gridgcn_oracle.c).  Parity
    unpinned."""
    fn = lib().gridgcn_oracle_gridify_occaware
    fn.restype = ctypes.c_int
    fn.argtypes = [_f32p, _i32p] + [ctypes.c_int] * 6 + [_f32p, _f32p, _i32p, ctypes.c_int, ctypes.c_int,
                                                         ctypes.c_ulonglong, _i32p, _f32p, _f32p, _f32p, _i32p]
    data, pd = _f(data)
    assert data.ndim == 3 and data.shape[2] == 4, "data should be (B, N, 4)"
    B, N, _ = data.shape
    npts, pn = _i(np.asarray(actual_numpoints).reshape(B))
    O, P = int(max_o_grid), int(max_p_grid)
    shift = _triple(coord_shift, np.float32)
    voxel = _triple(voxel_size, np.float32)
    grid = _triple(grid_size, np.int32)
    nebidx = np.empty((B, O, P), np.int32)
    nebmsk = np.empty((B, O, P), np.float32)
    cent = np.empty((B, O, 4), np.float32)
    centmsk = np.empty((B, O), np.float32)
    centnum = np.empty((B, 1), np.int32)
    rc = fn(pd, pn, B, N, O, P, int(kernel_size), int(loc), shift.ctypes.data_as(_f32p),
            voxel.ctypes.data_as(_f32p), grid.ctypes.data_as(_i32p), 1 if knn_query else 0,
            1 if dist_fma else 0, int(seed), nebidx.ctypes.data_as(_i32p),
            nebmsk.ctypes.data_as(_f32p), cent.ctypes.data_as(_f32p), centmsk.ctypes.data_as(_f32p),
            centnum.ctypes.data_as(_i32p))
    if rc != 0:
        raise ValueError("oracle rejected the arguments (rc=%d)" % rc)
    return nebidx, nebmsk, cent, centmsk, centnum


def gridify_up(downdata, updata, down_actual_numpoints, up_actual_numpoints, *, max_p_grid,
               max_o_grid, kernel_size, coord_shift=(0, 0, 0), voxel_size=(1, 1, 1),
               grid_size=(1, 1, 1)):
    """GridifyUp (gridify_up-inl.h:93-119, gridify_up.cu:102-225)."""
    downdata, pdd = _f(downdata)
    updata, pud = _f(updata)
    B, N, _ = downdata.shape
    O, P = int(max_o_grid), int(max_p_grid)
    assert updata.shape == (B, O, 4), "updata should be (B, max_o_grid, 4)"
    dn, pdn = _i(np.asarray(down_actual_numpoints).reshape(B))
    un, pun = _i(np.asarray(up_actual_numpoints).reshape(B))
    shift = _triple(coord_shift, np.float32)
    voxel = _triple(voxel_size, np.float32)
    grid = _triple(grid_size, np.int32)
    nebidx = np.empty((B, O, P), np.int32)
    nebmsk = np.empty((B, O, P), np.float32)
    fn = lib().gridgcn_oracle_gridify_up
    fn.restype = ctypes.c_int
    rc = fn(pdd, pud, pdn, pun, B, N, O, P, int(kernel_size), shift.ctypes.data_as(_f32p),
            voxel.ctypes.data_as(_f32p), grid.ctypes.data_as(_i32p),
            nebidx.ctypes.data_as(_i32p), nebmsk.ctypes.data_as(_f32p))
    if rc != 0:
        raise ValueError("oracle rejected the arguments (rc=%d)" % rc)
    return nebidx, nebmsk


def _knn(unknown, known, downnum, upnum, k, radius, dist_fma):
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    assert unknown.ndim == 3 and unknown.shape[2] == 3, "Last dim of unknown should be 3"
    assert known.ndim == 3 and known.shape[2] == 3, "Last dim of known should be 3"
    B, n, _ = unknown.shape
    m = known.shape[1]
    dn, pdn = _i(np.asarray(downnum).reshape(B))
    un, pun = _i(np.asarray(upnum).reshape(B))
    idx = np.empty((B, n, int(k)), np.int32)
    L = lib()
    if radius is None:
        L.gridgcn_oracle_knn.restype = ctypes.c_int
        rc = L.gridgcn_oracle_knn(pu, pk, pdn, pun, B, n, m, int(k), 1 if dist_fma else 0,
                                  idx.ctypes.data_as(_i32p))
    else:
        L.gridgcn_oracle_ball_knn.restype = ctypes.c_int
        L.gridgcn_oracle_ball_knn.argtypes = [_f32p, _f32p, _i32p, _i32p, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_float, ctypes.c_int, _i32p]
        rc = L.gridgcn_oracle_ball_knn(pu, pk, pdn, pun, B, n, m, int(k), float(radius),
                                       1 if dist_fma else 0, idx.ctypes.data_as(_i32p))
    if rc != 0:
        raise ValueError("oracle rejected the arguments (rc=%d)" % rc)
    return idx


def knn(unknown, known, downnum, upnum, *, k=3, dist_fma=False):
    """contrib.KNN (k_nn.cc:14-65, k_nn-inl.h:40-92)."""
    return _knn(unknown, known, downnum, upnum, k, None, dist_fma)


def ball_knn(unknown, known, downnum, upnum, *, k=3, radius=0.1, dist_fma=False):
    """contrib.BallKNN (ball_k_nn.cc:14-65, ball_k_nn-inl.h:43-95)."""
    return _knn(unknown, known, downnum, upnum, k, radius, dist_fma)
