"""GPU: the hand-written tcgen05 / TMEM primitives against a float64 matmul (descriptor encodings,
K-major no-swizzle operand layout, TMEM addressing, 3-term tf32 split)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(16, 8), (32, 32), (64, 16), (128, 64), (256, 64), (48, 24), (16, 128)])
@pytest.mark.parametrize("nsplit", [1, 3])
def test_tc_gemm_selftest(gg, cuda_dev, N, K, nsplit):
    L = gg._lib.lib()
    rng = np.random.default_rng(N * 1000 + K)
    A = rng.normal(size=(128, K)).astype(np.float32)
    B = rng.normal(size=(N, K)).astype(np.float32)
    a, b = torch.from_numpy(A).to(cuda_dev), torch.from_numpy(B).to(cuda_dev)
    d = torch.full((128, N), float("nan"), dtype=torch.float32, device=cuda_dev)
    rc = L.gridgcn_debug_tc_gemm(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, nsplit,
                                 torch.cuda.current_stream(cuda_dev).cuda_stream)
    gg._lib.check(rc, "gridgcn_debug_tc_gemm")
    torch.cuda.synchronize()
    want = A.astype(np.float64) @ B.astype(np.float64).T
    got = d.cpu().numpy().astype(np.float64)
    scale = np.sqrt(K)  # typical magnitude of an entry
    err = np.abs(got - want).max() / scale
    tol = 2e-3 if nsplit == 1 else 1e-5  # tf32: 2^-11 per operand; split: fp32-class
    print("N=%d K=%d nsplit=%d err=%.3g" % (N, K, nsplit, err))
    assert err < tol, "N=%d K=%d nsplit=%d: max err / sqrt(K) = %.3g" % (N, K, nsplit, err)


def test_tf32_operands_are_truncated(gg, cuda_dev):
    """The 3-pass split (tc_common.cuh split_op) relies on tcgen05 kind::tf32 TRUNCATING the low 13
    mantissa bits of fp32 operands.  nsplit == 0 feeds the self-test kernel raw fp32 bits."""
    L = gg._lib.lib()
    rng = np.random.default_rng(0)
    N, K = 64, 8
    A = rng.normal(size=(128, K)).astype(np.float32)
    B = rng.normal(size=(N, K)).astype(np.float32)
    a, b = torch.from_numpy(A).to(cuda_dev), torch.from_numpy(B).to(cuda_dev)
    d = torch.zeros((128, N), dtype=torch.float32, device=cuda_dev)
    gg._lib.check(L.gridgcn_debug_tc_gemm(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, 0,
                                          torch.cuda.current_stream(cuda_dev).cuda_stream), "tc_gemm")
    torch.cuda.synchronize()
    got = d.cpu().numpy().astype(np.float64)
    trunc = lambda x: (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    want = trunc(A).astype(np.float64) @ trunc(B).astype(np.float64).T
    exact = A.astype(np.float64) @ B.astype(np.float64).T
    assert np.abs(got - want).max() < 1e-5
    assert np.abs(got - exact).max() > 1e-4  # i.e. it really is tf32, not fp32


def test_xorwow_matches_device_curand(gg, cuda_dev):
    """The strict K2 reservoir and the coverage-aware sampling draw from the library's own XORWOW; the reference
    draws from cuRAND (curand_init(seed, 0, 0) + curand_uniform, gridify.cu:260-261).  Bit equality of the first
    uniform for small, large, negative-int-widened and random 64-bit seeds -- and of the CPU oracle's XORWOW."""
    import ctypes
    from oracle import oracle
    rng = np.random.default_rng(0)
    seeds = np.concatenate([np.arange(0, 4096, dtype=np.uint64),
                            np.array([2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1, 2 ** 32, 2 ** 63, 2 ** 64 - 1], np.uint64),
                            (np.arange(-2000, 0, dtype=np.int64)).astype(np.uint64),  # (long long)(int) seeds < 0
                            rng.integers(0, 2 ** 63, size=20000, dtype=np.uint64) * np.uint64(2) + np.uint64(1)])
    s = torch.from_numpy(seeds.view(np.int64)).to(cuda_dev)
    ours = torch.empty(len(seeds), dtype=torch.float32, device=cuda_dev)
    theirs = torch.empty_like(ours)
    rc = gg._lib.lib().gridgcn_debug_curand_first_uniform(s.data_ptr(), len(seeds), ours.data_ptr(),
                                                          theirs.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(ours, theirs)
    assert float(theirs.min()) > 0.0 and float(theirs.max()) <= 1.0
    # the CPU oracle's XORWOW (gridgcn_oracle.c) through its exported probe
    L = oracle.lib()
    L.gridgcn_oracle_xorwow_first_uniform.restype = ctypes.c_float
    L.gridgcn_oracle_xorwow_first_uniform.argtypes = [ctypes.c_ulonglong]
    host = np.array([L.gridgcn_oracle_xorwow_first_uniform(int(x)) for x in seeds[:6200]], np.float32)
    assert np.array_equal(host, theirs.cpu().numpy()[:6200])
