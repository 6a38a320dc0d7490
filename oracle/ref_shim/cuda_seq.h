// cuda_seq.h -- TEST INFRASTRUCTURE.  Lets the reference's CUDA kernel BODIES (gridifyop/gridify.cu,
// gridifyknn.cu, gridify_up.cu, compiled from where they lie -- see the Makefile rule of
// _ref/libgridify_ref.so) run on the host, one CUDA thread after the other in ascending global index:
// the canonical schedule of SURVEY.md s8(c).  Nothing here is reference code: it only supplies the
// CUDA built-ins the kernel bodies use (thread indices, the atomics -- trivially atomic in a sequential
// schedule -- and curand_init / curand_uniform as XORWOW per /usr/local/cuda/include/curand_kernel.h:772-798,
// 863-874, curand_uniform.h:69-72; the same generator is pinned against the device cuRAND on the GPU box,
// tests/test_gpu_tc.py::test_xorwow_matches_device_curand).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define ndim 3        /* gridify.cu:15 */
#define data_ndim 4   /* gridify.cu:16 */

struct seq_dim3 {
    unsigned x, y, z;
};
static seq_dim3 blockIdx = {0, 0, 0}, blockDim = {1024, 1, 1}, threadIdx = {0, 0, 0};

using std::max;
using std::min;

static inline int atomicAdd(int *addr, int v) {
    int old = *addr;
    *addr = old + v;
    return old;
}
static inline float atomicAdd(float *addr, float v) {
    float old = *addr;
    *addr = old + v;
    return old;
}
static inline int atomicCAS(int *addr, int compare, int val) {
    int old = *addr;
    if (old == compare) *addr = val;
    return old;
}

struct curandState {
    uint32_t d, v[5];
};
static inline void curand_init(unsigned long long seed, unsigned long long subsequence, unsigned long long offset,
                               curandState *s) {
    if (subsequence != 0 || offset != 0) abort();  // the reference only uses (seed, 0, 0): no skip-ahead
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    s->d = 6615241u + t1 + t0;
    s->v[0] = 123456789u + t0;
    s->v[1] = 362436069u ^ t0;
    s->v[2] = 521288629u + t1;
    s->v[3] = 88675123u ^ t1;
    s->v[4] = 5783321u + t0;
}
static inline float curand_uniform(curandState *s) {
    uint32_t t = (s->v[0] ^ (s->v[0] >> 2));
    s->v[0] = s->v[1];
    s->v[1] = s->v[2];
    s->v[2] = s->v[3];
    s->v[3] = s->v[4];
    s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
    s->d += 362437u;
    return fmaf((float)(s->v[4] + s->d), 2.3283064e-10f, 2.3283064e-10f / 2.0f);  // one FFMA on the device
}

// "Launch": every thread of a 1-D grid of 1024-thread blocks (mshadow::cuda::kMaxThreadsPerBlock), in order.
template <class F>
static inline void seq_launch(long long total_threads, F body) {
    const long long blocks = (total_threads + 1023) / 1024;
    blockDim.x = 1024;
    for (long long b = 0; b < blocks; b++)
        for (unsigned t = 0; t < 1024; t++) {
            blockIdx.x = (unsigned)b;
            threadIdx.x = t;
            body();
        }
}
