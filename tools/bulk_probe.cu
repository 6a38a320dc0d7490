// Hardware probe (not part of the library): sustained rate of cp.async.bulk global->shared streaming when every
// SM walks the SAME buffer (the weight-streaming pattern of the GridConv kernels), as a function of slice size
// and of the number of copies kept in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/bulk_probe tools/bulk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../grid-gcn_b200/csrc/tc_common.cuh"
using namespace gg;

__global__ void __launch_bounds__(32) probe(const float *src, size_t total_bytes, int slice_bytes, int depth,
                                            int rotate, long long *cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; i++) tc::mbar_init(&full[i], 1);
        tc::mbar_init_fence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n = (int)(total_bytes / slice_bytes);
        const int start = rotate ? (blockIdx.x * 7) % n : 0;
        long long t0 = clock64();
        int issued = 0;
        for (; issued < depth && issued < n; issued++) {
            tc::mbar_expect_tx(&full[issued], slice_bytes);
            tc::bulk_g2s(smem + (size_t)issued * slice_bytes,
                         (const uint8_t *)src + (size_t)((start + issued) % n) * slice_bytes, slice_bytes, &full[issued]);
        }
        for (int c = 0; c < n; c++) {
            const int slot = c % depth;
            tc::mbar_wait(&full[slot], (c / depth) & 1);
            if (issued < n) {
                tc::mbar_expect_tx(&full[slot], slice_bytes);
                tc::bulk_g2s(smem + (size_t)slot * slice_bytes,
                             (const uint8_t *)src + (size_t)((start + issued) % n) * slice_bytes, slice_bytes, &full[slot]);
                issued++;
            }
        }
        cyc[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    const size_t total = 2u << 20;  // 2 MB "weight set"
    float *src; long long *cyc;
    cudaMalloc(&src, total); cudaMemset(src, 0, total);
    cudaMallocManaged(&cyc, 1024 * 8);
    float *flush; cudaMalloc(&flush, 256u << 20);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int blocks : {1, 148}) for (int rotate : {0, 1}) for (int slice : {8192, 32768}) for (int depth : {1, 2, 4, 6}) {
        if (blocks == 1 && rotate) continue;
        if ((size_t)slice * depth > 196 * 1024) continue;
        for (int warm = 0; warm < 2; warm++) {  // warm=0: cold L2 (flushed), warm=1: L2 resident
            if (warm == 0) cudaMemset(flush, 1, 256u << 20);
            probe<<<blocks, 32, slice * depth + 1024>>>(src, total, slice, depth, rotate, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long mx = 0; for (int i = 0; i < blocks; i++) mx = cyc[i] > mx ? cyc[i] : mx;
            double us = mx / (clk * 1e-3);
            printf("blocks=%3d rotate=%d slice=%5d depth=%d %s: %8.1f us  per-SM %6.1f GB/s  aggregate %7.1f GB/s\n", blocks,
                   rotate, slice, depth, warm ? "warm" : "cold", us, total / us * 1e-3, total * (double)blocks / us * 1e-3);
        }
    }
    return 0;
}
