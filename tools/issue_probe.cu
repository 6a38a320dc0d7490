// Hardware probe (not part of the library): is the ~71.5-cycle minimum per tcgen05.mma (tools/chain_probe.cu) a
// limit of the ISSUING THREAD or of the tensor pipe?  W warps issue 64 MMAs each, concurrently, into their own
// accumulators; the kernel reports the cycles until all of them have retired.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/issue_probe tools/issue_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../grid-gcn_b200/csrc/tc_common.cuh"
using namespace gg;

__global__ void __launch_bounds__(256) probe(long long *cyc, int M, int N, int W) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t lbo_b = (uint32_t)N * 16, lbo_a = 128 * 16;
    uint8_t *b = smem, *a_s = smem + 2 * lbo_b;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { for (int i = 0; i < 8; i++) tc::mbar_init(&bar[i], 1); tc::mbar_init_fence(); }
    for (int e = tid; e < (2 * (int)lbo_b + 2 * (int)lbo_a) / 4; e += 256) reinterpret_cast<float *>(smem)[e] = 1.f;
    tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    for (int rep = 0; rep < 2; rep++) {
        __syncthreads();
        const long long t0 = clock64();
        if (warp < W && lane == 0) {
            const uint32_t idesc = tc::make_idesc_tf32(M, N);
            const uint64_t bd = tc::make_sdesc(tc::smem_u32(b), lbo_b), ad = tc::make_sdesc(tc::smem_u32(a_s), lbo_a);
            const uint32_t d = tmem + (uint32_t)(warp * 64);
            for (int i = 0; i < 64; i++) tc::mma_tf32(d, ad, bd, idesc, 1);
            tc::mma_commit(&bar[warp]);
        }
        for (int w = 0; w < W; w++) tc::mbar_wait(&bar[w], rep & 1);
        tc::fence_after_sync();
        if (tid == 0) cyc[rep] = clock64() - t0;
    }
    tc::fence_before_sync(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
    long long *cyc;
    cudaMallocManaged(&cyc, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int M : {64, 128})
        for (int N : {32, 64})
            for (int W : {1, 2, 4, 8}) {
                if (M == 64 && N == 32) continue;
                probe<<<1, 256, 2 * N * 16 + 2 * 128 * 16 + 1024>>>(cyc, M, N, W);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("M=%d N=%d W=%d: %s\n", M, N, W, cudaGetErrorString(e)); return 1; }
                printf("M=%3d N=%3d issuers=%d: %6lld cycles for %d MMAs = %.1f cycles per MMA overall (math floor %d)\n", M, N, W,
                       cyc[1], 64 * W, cyc[1] / (64.0 * W), 128 * N / 256);
            }
    return 0;
}
