"""ctypes binding of libgridgcn_b200.so (the C-ABI declared in include/gridgcn_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or was not built, loading raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GRIDGCN_B200_LIB: developer override to A/B an experimental build of the same C-ABI (tools/)
LIB_PATH = os.environ.get("GRIDGCN_B200_LIB") or os.path.join(_HERE, "libgridgcn_b200.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_sz = ctypes.c_size_t
_f3 = ctypes.POINTER(ctypes.c_float)
_i3 = ctypes.POINTER(ctypes.c_int)

MAX_STAGES = 8


class MlpDesc(ctypes.Structure):
    """gridgcn_mlp_t"""
    _fields_ = [("n_feat_stages", _i), ("attfdim", _i), ("feat_in", _i),
                ("widths", _i * MAX_STAGES), ("weight", _vp * MAX_STAGES),
                ("bias", _vp * MAX_STAGES), ("pre_relu", _i), ("n_att_stages", _i), ("localfdim", _i),
                ("att_full", _i)]


_GRIDIFY_ARGS = [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f3, _f3, _i3, _i, _vp, _vp, _vp, _vp, _vp,
                 _vp, _sz, _vp]

# name -> (restype, argtypes); every symbol include/gridgcn_b200.h declares
SIGNATURES = {
    "gridgcn_abi_version": (_i, []),
    "gridgcn_strerror": (ctypes.c_char_p, [_i]),
    "gridgcn_gridify_workspace_bytes": (_sz, [_i, _i, _i, _i3]),
    "gridgcn_gridify_fwd": (_i, _GRIDIFY_ARGS),
    "gridgcn_gridify_knn_fwd": (_i, _GRIDIFY_ARGS),
    "gridgcn_gridify_occaware_workspace_bytes": (_sz, [_i, _i, _i, _i, _i3]),
    "gridgcn_gridify_occaware_fwd": (_i, _GRIDIFY_ARGS[:13] + [ctypes.c_ulonglong] + _GRIDIFY_ARGS[13:]),
    "gridgcn_gridify_up_workspace_bytes": (_sz, [_i, _i, _i3]),
    "gridgcn_gridify_up_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f3, _f3, _i3, _vp, _vp,
                                    _vp, _sz, _vp]),
    "gridgcn_knn_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "gridgcn_ball_knn_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp, _vp]),
    "gridgcn_rowmlp_fwd": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp,
                                ctypes.c_longlong, _vp]),
    "gridgcn_rowmlp_tc_fwd": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp,
                                   ctypes.c_longlong, _vp]),
    # training-mode block kernels (csrc/train_ops.cu)
    "gridgcn_train_edge_rows": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "gridgcn_train_col_sums": (_i, [_vp, _vp, ctypes.c_longlong, _i, _vp, _vp, _vp]),
    "gridgcn_train_bn_finalize": (_i, [_vp, _vp, ctypes.c_longlong, _i, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gridgcn_train_bn_relu_fwd": (_i, [_vp, ctypes.c_longlong, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gridgcn_train_relu_bwd_xhat": (_i, [_vp, _vp, _vp, ctypes.c_longlong, _i, _vp, _vp, _vp, _vp, _vp]),
    "gridgcn_train_bn_bwd": (_i, [_vp, _vp, ctypes.c_longlong, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gridgcn_train_pool_fwd": (_i, [_vp, _vp, ctypes.c_longlong, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "gridgcn_train_pool_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, ctypes.c_longlong, _i, _i, _vp, _vp, _vp]),
    "gridgcn_train_wgrad": (_i, [_vp, _i, _vp, _i, _i, _vp, _i, _i, ctypes.c_longlong, _vp, _vp, _vp]),
    "gridgcn_train_scatter_add": (_i, [_vp, _vp, ctypes.c_longlong, _i, _i, _vp, _vp]),
    "gridgcn_debug_tc_gemm": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "gridgcn_debug_curand_first_uniform": (_i, [_vp, _i, _vp, _vp, _vp]),
    "gridgcn_debug_phase_buffer": (None, [_vp]),
    "gridgcn_gridconv_packed_bytes": (_sz, [ctypes.POINTER(MlpDesc), _i]),
    "gridgcn_gridconv_pack": (_i, [ctypes.POINTER(MlpDesc), _i, _vp, _sz, _vp]),
    "gridgcn_gridconv_workspace_bytes": (_sz, [ctypes.POINTER(MlpDesc), _i, _i, _i]),
    "gridgcn_gridconv_fp32_scratch_bytes": (_sz, [ctypes.POINTER(MlpDesc), _i, _i]),
    "gridgcn_gridconv_edge_workspace_bytes": (_sz, [ctypes.POINTER(MlpDesc), _i, _i, _i, _i]),
    "gridgcn_gridconv_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, ctypes.POINTER(MlpDesc),
                                  _i, _vp, _vp, _sz, _vp, _vp]),
}

_lib = None


class GridGcnError(RuntimeError):
    """Raised when the C-ABI returns a non-zero code (the reference would LOG(FATAL) /
    raise MXNetError, gridify.cu:386, gridify-inl.h:174-182)."""


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GridGcnError(
                "libgridgcn_b200.so is missing (%s). Build it with __graft_entry__.build() or "
                "`python grid-gcn_b200/build.py`; there is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if L.gridgcn_abi_version() != 2:
            raise GridGcnError("libgridgcn_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().gridgcn_strerror(code).decode()
        raise GridGcnError("%s failed: %s (code %d)" % (what, msg, code))


def triple_f(v):
    v = list(v) if hasattr(v, "__len__") else [v] * 3
    if len(v) != 3:
        raise ValueError("expected 3 values, got %r" % (v,))
    return (ctypes.c_float * 3)(*[float(x) for x in v])


def triple_i(v):
    v = list(v) if hasattr(v, "__len__") else [v] * 3
    if len(v) != 3:
        raise ValueError("expected 3 values, got %r" % (v,))
    return (ctypes.c_int * 3)(*[int(x) for x in v])
