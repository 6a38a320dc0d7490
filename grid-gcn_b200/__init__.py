"""gridgcn_b200 -- B200-native (sm_100a) drop-in for Grid-GCN's grid query + GridConv hot path.

Only the path named by BASELINE.json:north_star lives here: the five native
operators and the fused GridConv block, as CUDA kernels behind a C-ABI library
(``include/gridgcn_b200.h``), with a Python host side that mirrors the reference's operator
signatures.  There is no CPU fallback: every op raises if the CUDA library is missing or the
tensors are not on a CUDA device.
"""
from . import synth  # noqa: F401  host-side input generator, numpy only
from . import _lib  # noqa: F401  ctypes binding (loads the library lazily)
from .ops import Gridify, GridifyKNN, GridifyUp, Gridify_occaware, GridifyOccaware, contrib  # noqa: F401
from .gridconv import (GridConv, GridConvUp, SegHead, sub_g_update, fold_bn, init_layer,  # noqa: F401
                       init_up_layer, features_nco, rowmlp)
from . import stack, shard, train  # noqa: F401
from .build import build  # noqa: F401

__all__ = ["Gridify", "GridifyKNN", "GridifyUp", "Gridify_occaware", "GridifyOccaware", "contrib", "GridConv", "sub_g_update", "fold_bn",
           "init_layer", "features_nco", "stack", "shard", "synth", "build"]
