// Hardware probe (not part of the library): throughput of redux.sync.max.u32 against FMNMX / SHFL, one warp and
// four warps per SM sub-partition, for the cross-lane max pool of a non-transposed final stage.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/redux_probe tools/redux_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned *out, long long *cyc, int mode) {
    unsigned v[8];
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 2654435761u + i;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (mode == 0) v[i] = __reduce_max_sync(0xffffffffu, v[i] + it);
            else if (mode == 1) v[i] = max(v[i], __shfl_xor_sync(0xffffffffu, v[i], 16)) + it;
            else v[i] = max(v[i] + it, v[(i + 1) & 7]);
        }
    }
    const long long t1 = clock64();
    unsigned s = 0;
    for (int i = 0; i < 8; i++) s += v[i];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[mode] = t1 - t0;
}
int main() {
    unsigned *out; long long *cyc;
    cudaMallocManaged(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 64);
    for (int threads : {32, 128, 512})
        for (int mode = 0; mode < 3; mode++) {
            probe<<<148, threads>>>(out, cyc, mode);
            cudaDeviceSynchronize();
            printf("threads/CTA %4d  %s: %.2f cycles per warp-instruction (%.2f per SM)\n", threads,
                   mode == 0 ? "REDUX.MAX" : mode == 1 ? "SHFL+max " : "IADD+max ", cyc[mode] / 2048.0, cyc[mode] / 2048.0 / (threads / 32));
        }
    return 0;
}
