// Brute-force KNN / BallKNN (sm_100a): one thread per query row, known points staged through
// shared memory tile by tile, top-k kept sorted in registers.
//
// Replaces KNNKernel::Map (reference gridifyop/k_nn-inl.h:40-92, which heap-allocates best[] per
// thread with device-side new[] and streams all known points from global memory per thread) and
// BallKNNKernel::Map (ball_k_nn-inl.h:43-95).  Same visiting order (k = 0..downnum-1) and the
// same strict-< insertion, so ties resolve to the lowest known index.
#pragma once
#include "common.cuh"

namespace gg {

constexpr int kKnnThreads = 256;
constexpr int kKnnTile = 1024;  // known points per shared-memory tile (12 KB)

template <int KMAX, bool BALL>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(const float *__restrict__ unknown, const float *__restrict__ known,
           const int *__restrict__ downnum, const int *__restrict__ upnum, int n, int m, int k,
           float r2, int fma, int *__restrict__ idx) {
    __shared__ float tile[kKnnTile * 3];
    const int b = blockIdx.y;
    const int q = blockIdx.x * kKnnThreads + threadIdx.x;
    const int dn = min(max(downnum[b], 0), m), un = upnum[b];
    const bool active = q < n && q < un;
    const float *kn = known + (size_t)b * m * 3;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (q < n) {
        const float *u = unknown + ((size_t)b * n + q) * 3;
        ux = u[0];
        uy = u[1];
        uz = u[2];
    }
    float best[KMAX];
    int besti[KMAX];
#pragma unroll
    for (int l = 0; l < KMAX; l++) {
        best[l] = 3.402823466e+38f;  // FLT_MAX, k_nn-inl.h:65
        besti[l] = BALL ? -1 : 0;    // ball_k_nn-inl.h:69; KNN: oracle definition 0
    }
    for (int base = 0; base < dn; base += kKnnTile) {
        const int cnt = min(kKnnTile, dn - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 3; i += kKnnThreads) tile[i] = kn[(size_t)base * 3 + i];
        __syncthreads();
        if (!active) continue;
        for (int j = 0; j < cnt; j++) {
            float d = dist2(ux, uy, uz, tile[j * 3], tile[j * 3 + 1], tile[j * 3 + 2], fma);
            if (BALL && d > r2) continue;  // ball_k_nn-inl.h:77
            if (KMAX <= 8) {
                // Registers only: keep the best KMAX sorted (its first k entries equal the
                // reference's best-k list, same strict-< / lowest-index tie rule), all indices static.
                if (!(d < best[KMAX - 1])) continue;
                int pos = 0;  // insert before the first entry greater than d
#pragma unroll
                for (int l = 0; l < KMAX; l++) pos += !(d < best[l]) ? 1 : 0;
#pragma unroll
                for (int l = KMAX - 1; l >= 1; l--) {
                    if (l > pos) {
                        best[l] = best[l - 1];
                        besti[l] = besti[l - 1];
                    } else if (l == pos) {
                        best[l] = d;
                        besti[l] = base + j;
                    }
                }
                if (pos == 0) {
                    best[0] = d;
                    besti[0] = base + j;
                }
            } else {
                for (int l = 0; l < k; l++) {  // k_nn-inl.h:74-85 verbatim semantics
                    if (d < best[l]) {
                        for (int jj = k - 1; jj > l; jj--) {
                            best[jj] = best[jj - 1];
                            besti[jj] = besti[jj - 1];
                        }
                        best[l] = d;
                        besti[l] = base + j;
                        break;
                    }
                }
            }
        }
    }
    if (q < n) {
        int *out = idx + ((size_t)b * n + q) * k;
        if (KMAX <= 8) {
#pragma unroll
            for (int l = 0; l < KMAX; l++)
                if (l < k) out[l] = active ? besti[l] : 0;  // rows >= upnum: 0
        } else {
            for (int l = 0; l < k; l++) out[l] = active ? besti[l] : 0;
        }
    }
}

}  // namespace gg
