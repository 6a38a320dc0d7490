// Hardware probe (not part of the library): where does a cta_group::1 M=64 tcgen05.mma put its 64 accumulator
// rows in TMEM, which lane offsets are accepted, and how long do M=64 / M=128 MMAs take back to back?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/m64_probe tools/m64_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../grid-gcn_b200/csrc/tc_common.cuh"
using namespace gg;

__global__ void __launch_bounds__(128) probe(float *D, long long *cyc, int lane_off, int N) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = 8;
    const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)N * 16;
    uint8_t *a128 = smem, *a64 = a128 + (K / 4) * lbo_a, *b = a64 + (K / 4) * lbo_a;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_init_fence(); }
    for (int e = tid; e < 128 * K; e += 128) {
        int r = e / K, k = e % K;
        *reinterpret_cast<float *>(a128 + tc::kmajor_off(r, k, lbo_a)) = k == 0 ? -(float)(r + 1) : 0.f;
        *reinterpret_cast<float *>(a64 + tc::kmajor_off(r, k, lbo_a)) = (k == 0 && r < 64) ? (float)(r + 1) : 0.f;
    }
    for (int e = tid; e < N * K; e += 128) {
        int r = e / K, k = e % K;
        *reinterpret_cast<float *>(b + tc::kmajor_off(r, k, lbo_b)) = k == 0 ? 1.f : 0.f;
    }
    tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;
    if (tid == 0) {
        tc::mma_tf32(tmem_d, tc::make_sdesc(tc::smem_u32(a128), lbo_a), tc::make_sdesc(tc::smem_u32(b), lbo_b),
                     tc::make_idesc_tf32(128, N), 0);
        tc::mma_tf32(tmem_d + ((uint32_t)lane_off << 16), tc::make_sdesc(tc::smem_u32(a64), lbo_a),
                     tc::make_sdesc(tc::smem_u32(b), lbo_b), tc::make_idesc_tf32(64, N), 0);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    {
        uint32_t v[16];
        tc::tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16), v);
        tc::tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[tid * 16 + j] = __uint_as_float(v[j]);
    }
    tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    // timing: 64 back-to-back MMAs of M=128 then of M=64 (N columns), issue -> retired
    for (int m = 0; m < 2; m++) {
        long long t0 = 0;
        if (tid == 0) {
            t0 = clock64();
            const uint32_t idesc = tc::make_idesc_tf32(m == 0 ? 128 : 64, N);
            for (int i = 0; i < 64; i++)
                tc::mma_tf32(tmem_d, tc::make_sdesc(tc::smem_u32(m == 0 ? a128 : a64), lbo_a),
                             tc::make_sdesc(tc::smem_u32(b), lbo_b), idesc, 1);
            tc::mma_commit(&bar);
        }
        tc::mbar_wait(&bar, (m + 1) & 1);
        tc::fence_after_sync();
        if (tid == 0) cyc[m] = clock64() - t0;
        __syncthreads();
    }
    tc::fence_before_sync(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, 256);
}

int main() {
    float *D; long long *cyc;
    cudaMallocManaged(&D, 128 * 16 * 4); cudaMallocManaged(&cyc, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int offs[4] = {0, 16, 64, 32};
    for (int N : {16, 128}) {
        for (int lo : offs) {
            if (N == 128 && lo != 0) continue;
            cudaMemset(D, 0, 128 * 16 * 4);
            probe<<<1, 128, 2 * 2 * 128 * 16 + 2 * N * 16 + 1024>>>(D, cyc, lo, N);
            cudaError_t e = cudaDeviceSynchronize();
            printf("N=%d lane_off=%d: %s\n", N, lo, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
            for (int l = 0; l < 128; l++) { printf("%5.0f", D[l * 16]); if (l % 16 == 15) printf("\n"); }
            printf("64 MMAs: M=128 %lld cycles, M=64 %lld cycles\n", cyc[0], cyc[1]);
        }
    }
    return 0;
}
