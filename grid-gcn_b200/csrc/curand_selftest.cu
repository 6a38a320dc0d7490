// Pins the library's XORWOW (common.cuh: xorwow_first_uniform, used by the strict K2 reservoir and by the
// coverage-aware sampling) against the REAL device cuRAND the reference kernels call
// (gridify.cu:260-261: curand_init(seed, 0, 0, &state); curand_uniform(&state)): for every seed both
// values are written side by side and tests/test_gpu_tc.py demands bit equality.
#include <curand_kernel.h>

#include "../../include/gridgcn_b200.h"
#include "common.cuh"

namespace gg {

__global__ void curand_selftest_kernel(const unsigned long long *__restrict__ seeds, int n,
                                       float *__restrict__ ours, float *__restrict__ theirs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ours[i] = xorwow_first_uniform(seeds[i]);
    curandState st;
    curand_init(seeds[i], 0, 0, &st);
    theirs[i] = curand_uniform(&st);
}

}  // namespace gg

extern "C" int gridgcn_debug_curand_first_uniform(const unsigned long long *seeds, int n, float *ours,
                                                  float *theirs, void *stream) {
    if (!seeds || !ours || !theirs || n < 0) return GRIDGCN_EINVAL;
    if (n == 0) return 0;
    gg::curand_selftest_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(seeds, n, ours,
                                                                                              theirs);
    return (int)cudaGetLastError();
}
