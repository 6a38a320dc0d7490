"""The restatement (and, on the GPU box, the CUDA kernels) against the REFERENCE's own KNN / BallKNN
kernel bodies, compiled from /root/reference through oracle/ref_shim into oracle/_ref/libknn_ref.so
(`make -C oracle ref`).  These are the only two operators of the path with a CPU implementation in
the reference (k_nn.cc:59, ball_k_nn.cc:59).  Skipped when the library has not been built."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libknn_ref.so")


def _ref():
    if not os.path.exists(REF):
        from oracle import oracle
        try:
            oracle.build_ref()
        except Exception:
            pass
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libknn_ref.so not built (needs /root/reference)")
    L = ctypes.CDLL(REF)
    L.ref_ball_knn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_void_p]
    return L


def _case(seed, B, n, m):
    rng = np.random.default_rng(seed)
    unknown = rng.uniform(-1, 1, size=(B, n, 3)).astype(np.float32)
    known = rng.uniform(-1, 1, size=(B, m, 3)).astype(np.float32)
    known[0, m // 2:] = known[0, :m - m // 2]  # exact ties
    downnum = np.full((B, 1), m, np.int32)
    upnum = np.full((B, 1), n, np.int32)
    downnum[-1, 0], upnum[-1, 0] = m - m // 4, n - n // 3
    return unknown, known, downnum, upnum


def _run_ref(L, unknown, known, downnum, upnum, k, radius=None):
    B, n, _ = unknown.shape
    m = known.shape[1]
    idx = np.zeros((B, n, k), np.int32)  # rows >= upnum are left unwritten by the reference
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    if radius is None:
        L.ref_knn(p(unknown), p(known), p(downnum), p(upnum), B, n, m, k, p(idx))
    else:
        L.ref_ball_knn(p(unknown), p(known), p(downnum), p(upnum), B, n, m, k, radius, p(idx))
    return idx


def test_oracle_matches_reference_knn_bodies(oracle_mod):
    L = _ref()
    for seed, (B, n, m) in enumerate(((2, 300, 200), (1, 64, 1000), (3, 50, 24))):
        unknown, known, downnum, upnum = _case(seed, B, n, m)
        for k in (1, 3, 5, 8):
            assert np.array_equal(_run_ref(L, unknown, known, downnum, upnum, k),
                                  oracle_mod.knn(unknown, known, downnum, upnum, k=k)), (B, n, m, k)
        for k, r in ((5, 0.3), (3, 0.1), (6, 1.02)):
            assert np.array_equal(_run_ref(L, unknown, known, downnum, upnum, k, r),
                                  oracle_mod.ball_knn(unknown, known, downnum, upnum, k=k, radius=r)), (B, n, m, k)


@pytest.mark.gpu
def test_cuda_matches_reference_knn_bodies(gg, cuda_dev):
    import torch
    L = _ref()
    for seed, (B, n, m) in enumerate(((2, 1024, 256), (1, 8192, 1024))):
        unknown, known, downnum, upnum = _case(seed, B, n, m)
        args = [torch.from_numpy(a).to(cuda_dev) for a in (unknown, known, downnum, upnum)]
        for k in (3, 5):
            got = gg.contrib.KNN(*args, k=k).cpu().numpy()
            assert np.array_equal(got, _run_ref(L, unknown, known, downnum, upnum, k)), ("knn", k)
        for k, r in ((5, 0.34), (5, 1.02)):  # decoder radii, ggcn_models_g.py:204
            got = gg.contrib.BallKNN(*args, k=k, radius=r).cpu().numpy()
            assert np.array_equal(got, _run_ref(L, unknown, known, downnum, upnum, k, r)), ("ball", k)
