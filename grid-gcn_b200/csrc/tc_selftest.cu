// Self-test of the tcgen05 primitives (tc_common.cuh): one CTA computes D[128 x N] = A[128 x K] *
// B[N x K]^T from row-major fp32 inputs with kind::tf32 MMAs (1-term or 3-term split) and writes D
// back.  tests/test_gpu_tc.py compares it with a float64 matmul; it pins the descriptor encodings
// (instruction descriptor, shared-memory matrix descriptor, TMEM addressing) on real hardware.
#include "../../include/gridgcn_b200.h"
#include "tc_common.cuh"

namespace gg {

__global__ void __launch_bounds__(128)
tc_gemm_selftest_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D,
                        int N, int K, int nsplit) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)N * 16;  // panel = rows x 16 B
    const uint32_t a_bytes = (K / 4) * lbo_a, b_bytes = (K / 4) * lbo_b;
    uint8_t *a_hi = smem, *a_lo = a_hi + a_bytes, *b_hi = a_lo + a_bytes, *b_lo = b_hi + b_bytes;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_init_fence();
    }
    for (int e = tid; e < 128 * K; e += 128) {
        int r = e / K, k = e % K;
        float hi, lo;
        tc::split_tf32(A[e], hi, lo);
        if (nsplit == 0) hi = A[e];  // probe: unrounded fp32 bits
        uint32_t off = tc::kmajor_off(r, k, lbo_a);
        *reinterpret_cast<float *>(a_hi + off) = hi;
        *reinterpret_cast<float *>(a_lo + off) = lo;
    }
    for (int e = tid; e < N * K; e += 128) {
        int r = e / K, k = e % K;
        float hi, lo;
        tc::split_tf32(B[e], hi, lo);
        if (nsplit == 0) hi = B[e];
        uint32_t off = tc::kmajor_off(r, k, lbo_b);
        *reinterpret_cast<float *>(b_hi + off) = hi;
        *reinterpret_cast<float *>(b_lo + off) = lo;
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, N);
        uint32_t acc = 0;
        for (int ks = 0; ks < K / 8; ks++) {
            uint64_t ah = tc::make_sdesc(tc::smem_u32(a_hi) + ks * 2 * lbo_a, lbo_a);
            uint64_t bh = tc::make_sdesc(tc::smem_u32(b_hi) + ks * 2 * lbo_b, lbo_b);
            if (nsplit == 3) {  // small terms first
                uint64_t al = tc::make_sdesc(tc::smem_u32(a_lo) + ks * 2 * lbo_a, lbo_a);
                uint64_t bl = tc::make_sdesc(tc::smem_u32(b_lo) + ks * 2 * lbo_b, lbo_b);
                tc::mma_tf32(tmem_d, al, bh, idesc, acc);
                tc::mma_tf32(tmem_d, ah, bl, idesc, 1);
                acc = 1;
            }
            tc::mma_tf32(tmem_d, ah, bh, idesc, acc);
            acc = 1;
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j++) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, 256);
}

}  // namespace gg

// Debug / self-test entry point (not part of the operator ABI).
extern "C" int gridgcn_debug_tc_gemm(const float *A, const float *B, float *D, int N, int K, int nsplit,
                                     void *stream) {
    if (!A || !B || !D || N < 16 || N > 256 || (N & 15) || K < 8 || (K & 7)) return GRIDGCN_EINVAL;
    size_t smem = (size_t)2 * (K / 4) * (128 + N) * 16;
    if (smem > 200 * 1024) return GRIDGCN_ELIMIT;
    cudaError_t e = cudaFuncSetAttribute(gg::tc_gemm_selftest_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    gg::tc_gemm_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(A, B, D, N, K, nsplit);
    return (int)cudaGetLastError();
}
