// Hardware probe (not part of the library): how long does a tcgen05.mma kind::tf32 take when it accumulates
// into the SAME TMEM tile as its predecessor, and how much of that is hidden when c independent accumulators
// are interleaved?  64 MMAs issued back to back by one thread, round-robin over c tiles, issue -> retired.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/chain_probe tools/chain_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../grid-gcn_b200/csrc/tc_common.cuh"
using namespace gg;

__global__ void __launch_bounds__(128) probe(long long *cyc, int M, int N, int chains, int stride, int lane_alt) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t lbo_b = (uint32_t)N * 16, lbo_a = 128 * 16;
    uint8_t *b = smem, *a_s = smem + 2 * lbo_b;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_init_fence(); }
    for (int e = tid; e < (2 * (int)lbo_b + 2 * (int)lbo_a) / 4; e += 128) reinterpret_cast<float *>(smem)[e] = 1.f;
    tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    for (int rep = 0; rep < 2; rep++) {
        long long t0 = 0;
        if (tid == 0) {
            const uint32_t idesc = tc::make_idesc_tf32(M, N);
            const uint64_t bd = tc::make_sdesc(tc::smem_u32(b), lbo_b), ad = tc::make_sdesc(tc::smem_u32(a_s), lbo_a);
            t0 = clock64();
            int ch = 0;
            for (int i = 0; i < 64; i++) {
                uint32_t d = tmem + (uint32_t)((lane_alt ? (ch >> 1) : ch) * stride);
                if (lane_alt && (ch & 1)) d += 16u << 16;
                tc::mma_tf32(d, ad, bd, idesc, 1);
                if (++ch == chains) ch = 0;
            }
            tc::mma_commit(&bar);
        }
        tc::mbar_wait(&bar, rep & 1);
        tc::fence_after_sync();
        if (tid == 0) cyc[rep] = clock64() - t0;
        __syncthreads();
    }
    tc::fence_before_sync(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
    long long *cyc;
    cudaMallocManaged(&cyc, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    struct Cfg { int M, N, chains, stride, lane_alt; };
    const Cfg cfgs[] = {{128, 32, 1, 32, 0}, {128, 32, 2, 32, 0}, {128, 32, 4, 32, 0}, {128, 32, 8, 32, 0},
                        {128, 48, 1, 48, 0}, {128, 48, 2, 48, 0}, {128, 48, 4, 48, 0},
                        {64, 64, 1, 64, 0}, {64, 64, 2, 64, 1}, {64, 64, 2, 64, 0}, {64, 64, 4, 64, 1}, {64, 64, 4, 64, 0}, {64, 64, 8, 64, 1},
                        {64, 128, 1, 128, 0}, {64, 128, 2, 128, 1}, {64, 128, 4, 128, 1},
                        {64, 256, 1, 256, 0}, {64, 256, 2, 256, 1},
                        {128, 128, 1, 128, 0}, {128, 128, 2, 128, 0}, {128, 256, 1, 256, 0}, {128, 256, 2, 256, 0}};
    for (const Cfg &c : cfgs) {
        probe<<<1, 128, 2 * c.N * 16 + 2 * 128 * 16 + 1024>>>(cyc, c.M, c.N, c.chains, c.stride, c.lane_alt);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d N=%d: %s\n", c.M, c.N, cudaGetErrorString(e)); return 1; }
        printf("M=%3d N=%3d chains=%d lane_alt=%d: %6.1f cycles per MMA (math floor %d)\n", c.M, c.N, c.chains, c.lane_alt,
               cyc[1] / 64.0, 128 * c.N / 256);
    }
    return 0;
}
