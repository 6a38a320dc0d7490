"""CPU suite, part 2: the C-ABI library builds for sm_100a without a GPU, loads, and exports every
symbol include/gridgcn_b200.h declares; argument validation returns codes instead of launching."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gridgcn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gridgcn_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = _declared_symbols()
    for s in ("gridgcn_gridify_fwd", "gridgcn_gridify_knn_fwd", "gridgcn_gridify_up_fwd",
              "gridgcn_knn_fwd", "gridgcn_ball_knn_fwd", "gridgcn_gridconv_fwd"):
        assert s in syms


def test_library_exports_every_declared_symbol(gg):
    lib = ctypes.CDLL(gg._lib.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(lib, s), "libgridgcn_b200.so does not export %s" % s
    assert set(gg._lib.SIGNATURES) == set(_declared_symbols())
    assert gg._lib.lib().gridgcn_abi_version() == 2


def test_library_is_sm100a_native(gg):
    """The cubin inside the library targets sm_100a (no PTX-JIT fallback to an older arch)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", gg._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_library_uses_blackwell_tensor_and_copy_engines(gg):
    """SASS evidence (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, bulk copies ->
    UBLKCP): the GridConv kernels really are tcgen05 / TMEM / TMA-engine code, not mma.sync recompiles."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", gg._lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass and "WGMMA" not in sass


def test_argument_validation_without_gpu(gg):
    """Rejected arguments never reach a launch, so these calls are safe on a CPU-only box."""
    L = gg._lib.lib()
    g3, f3 = gg._lib.triple_i([40] * 3), gg._lib.triple_f([0.05] * 3)
    assert L.gridgcn_gridify_workspace_bytes(2, 1024, 1024, g3) > 0
    assert L.gridgcn_gridify_workspace_bytes(2, 1024, 1024, gg._lib.triple_i([128] * 3)) == 0
    null = None
    # null pointers
    rc = L.gridgcn_gridify_fwd(null, null, 1, 8, 4, 4, 3, 1, 1, f3, f3, g3, 0, null, null, null, null,
                               null, null, 0, null)
    assert rc == -1
    # even kernel size (coor_indx_b_origin undefined in the reference, gridify.cu:248)
    rc = L.gridgcn_gridify_fwd(16, 16, 1, 8, 4, 4, 2, 1, 1, f3, f3, g3, 0, 16, 16, 16, 16, 16, 16, 0, null)
    assert rc == -1
    # P > 128 (best[128], gridifyknn.cu:257)
    rc = L.gridgcn_gridify_knn_fwd(16, 16, 1, 8, 4, 129, 3, 1, 1, f3, f3, g3, 0, 16, 16, 16, 16, 16, 16,
                                   0, null)
    assert rc == -2
    # workspace too small
    rc = L.gridgcn_gridify_fwd(16, 16, 1, 8, 4, 4, 3, 1, 1, f3, f3, g3, 0, 16, 16, 16, 16, 16, 16, 8, null)
    assert rc == -3
    # BallKNN k > 6 (best[6], ball_k_nn-inl.h:63-64)
    assert L.gridgcn_ball_knn_fwd(16, 16, 16, 16, 1, 4, 4, 7, 0.5, 0, 16, null) == -2
    assert L.gridgcn_knn_fwd(16, 16, 16, 16, 1, 4, 4, 0, 0, 16, null) == -1
    assert b"range" in L.gridgcn_strerror(-2)


def test_ops_refuse_cpu_tensors(gg):
    import torch
    data = torch.zeros(1, 8, 4)
    num = torch.full((1, 1), 8, dtype=torch.int32)
    kw = dict(max_p_grid=4, max_o_grid=4, kernel_size=3, coord_shift=[1] * 3, voxel_size=[0.5] * 3,
              grid_size=[4] * 3)
    for fn in (gg.Gridify, gg.GridifyKNN):
        with pytest.raises(gg._lib.GridGcnError):
            fn(data, num, **kw)
    with pytest.raises(gg._lib.GridGcnError):
        gg.contrib.KNN(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3), num, num, k=3)


def test_missing_library_fails_loudly(gg, monkeypatch):
    monkeypatch.setattr(gg._lib, "_lib", None)
    monkeypatch.setattr(gg._lib, "LIB_PATH", "/nonexistent/libgridgcn_b200.so")
    with pytest.raises(gg._lib.GridGcnError):
        gg._lib.lib()


def test_mlp_description_validation_without_gpu(gg):
    """gridgcn_mlp_t: the classification-block fields are validated on the host; the tensor-core path takes them as an
    un-fused chain of row GEMMs (its own packed image + per-edge workspace), the fp32 path reports the activation
    scratch the wide layers need."""
    import ctypes
    L = gg._lib.lib()

    def desc(pt, att, cin, localfdim=0, att_full=0, attfdim=4):
        d = gg._lib.MlpDesc()
        d.n_feat_stages, d.attfdim, d.pre_relu = len(pt), attfdim, 1
        d.n_att_stages, d.localfdim, d.att_full = len(att), localfdim, att_full
        d.feat_in = 3 if cin == 0 else cin + localfdim
        for i, w in enumerate(list(pt) + list(att)):
            d.widths[i] = w
            d.weight[i] = 16  # never dereferenced: nothing is launched here
            d.bias[i] = 16
        return d

    seg = desc([64, 64, 128], [32, 128], 64, attfdim=10)
    assert L.gridgcn_gridconv_packed_bytes(ctypes.byref(seg), 64) > 0
    assert L.gridgcn_gridconv_fp32_scratch_bytes(ctypes.byref(seg), 64, 64) == 0
    cls = desc([128, 128, 256], [128, 256, 256], 128, localfdim=3, att_full=1)
    L.gridgcn_gridconv_edge_workspace_bytes.restype = ctypes.c_size_t
    assert L.gridgcn_gridconv_packed_bytes(ctypes.byref(cls), 128) > 0           # tensor cores: chain of row GEMMs
    assert L.gridgcn_gridconv_edge_workspace_bytes(ctypes.byref(cls), 4, 128, 128, 64) > 0
    assert L.gridgcn_gridconv_edge_workspace_bytes(ctypes.byref(seg), 4, 64, 128, 64) == 0  # fused kernels: none
    assert L.gridgcn_gridconv_fp32_scratch_bytes(ctypes.byref(cls), 128, 64) > 0  # too wide for shared memory
    assert L.gridgcn_gridconv_pack(ctypes.byref(cls), 128, None, 0, None) == -3  # GRIDGCN_EWORKSPACE: no buffer
    bad = desc([64, 128], [32, 64], 16, att_full=1)  # last attention width must equal the feature width
    assert L.gridgcn_gridconv_fp32_scratch_bytes(ctypes.byref(bad), 16, 8) == 0
    assert L.gridgcn_gridconv_fwd(16, 16, 16, 16, 1, 8, 16, 4, 8, ctypes.byref(bad), 0, None, None, 0, 16, None) == -1
