// Shared device helpers for the gridgcn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace gg {

// "Already configured on the current device?" for kernel attributes (cudaFuncSetAttribute is per device and
// the library may serve several devices and host threads of one process).  A lost race only repeats an
// idempotent call.
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0};
    __host__ bool done(int dev) const { return dev >= 0 && dev < 64 && ((mask.load(std::memory_order_acquire) >> dev) & 1ull); }
    __host__ void set(int dev) { if (dev >= 0 && dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release); }
};

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// Parameters of one Gridify-family call, passed by value to the kernels.
// Mirrors GridifyParam (reference gridifyop/gridify-inl.h:58-87).
struct GridParams {
    float shift[3];
    float voxel[3];
    float gridf[3];  // grid size as float: the reference compares against float (gridify.cu:135)
    int grid[3];
    int G;           // grid[0]*grid[1]*grid[2]
    int W;           // bitmap words = ceil(G/32)
    int B, N, O, P, ks, loc, flags;
};

// Per-cloud workspace, all offsets in 32-bit words from the cloud's base (each a multiple of 4).
//   bitmap[W]   occupancy bit per voxel           wordpfx[W]  exclusive popcount prefix per word
//   vend[V]     end offset of each occupied voxel's segment in `sorted` (V = min(N, G))
//   sorted[N]   point ids ordered by (voxel, id)  cent_lin[O] linear voxel index of each centre
//   cent_acc[4*O] barycentre sums (x*w, y*w, z*w, w) of each centre's voxel
//   key[N], tmp[N]  per-point scratch used only when the cloud does not fit in shared memory
//   firstmap[N/32], firstpfx[N/32]  first-occurrence bitmap and its prefix (grid_build_multi.cuh)
struct WsLayout {
    int bitmap, wordpfx, vend, sorted, cent_lin, cent_acc, key, tmp, firstmap, firstpfx;
    long long stride;  // words per cloud
};

__host__ __device__ inline int round4(long long x) { return (int)((x + 3) & ~3LL); }

__host__ inline WsLayout make_layout(int N, int O, int G) {
    WsLayout L;
    int W = (G + 31) / 32;
    int V = N < G ? N : G;
    long long off = 4;  // word 0 = nocc
    L.bitmap = (int)off;   off += round4(W);
    L.wordpfx = (int)off;  off += round4(W);
    L.vend = (int)off;     off += round4(V);
    L.sorted = (int)off;   off += round4(N);
    L.cent_lin = (int)off; off += round4(O);
    L.cent_acc = (int)off; off += round4(4LL * O);
    L.key = (int)off;      off += round4(N);
    L.tmp = (int)off;      off += round4(N);
    L.firstmap = (int)off; off += round4((N + 31) / 32);  // first-occurrence bits / prefix: multi-kernel build only
    L.firstpfx = (int)off; off += round4((N + 31) / 32);
    L.stride = off;
    return L;
}

// A.1 voxelise (reference gridify.cu:134-143): fp32 add, IEEE fp32 divide, floor, bounds test
// against the float grid size.  Returns the linear voxel index or -1.
__device__ __forceinline__ int voxel_of(float x, float y, float z, const GridParams &g) {
    float q0 = __fdiv_rn(__fadd_rn(x, g.shift[0]), g.voxel[0]);
    float q1 = __fdiv_rn(__fadd_rn(y, g.shift[1]), g.voxel[1]);
    float q2 = __fdiv_rn(__fadd_rn(z, g.shift[2]), g.voxel[2]);
    int c0 = (int)floorf(q0), c1 = (int)floorf(q1), c2 = (int)floorf(q2);
    if (c0 < 0 || (float)c0 >= g.gridf[0]) return -1;
    if (c1 < 0 || (float)c1 >= g.gridf[1]) return -1;
    if (c2 < 0 || (float)c2 >= g.gridf[2]) return -1;
    return c2 * (g.grid[0] * g.grid[1]) + c1 * g.grid[0] + c0;  // exact == float path for G < 2^24
}

// Squared distance in the reference's association order (gridifyknn.cu:287, k_nn-inl.h:73).
// fma == 0: every product and sum rounded (SURVEY s8c rule 10); fma == 1: the contraction of the
// reference's shipped cubin, fma(dz,dz,fma(dy,dy,dx*dx)).
__device__ __forceinline__ float dist2(float ux, float uy, float uz, float x, float y, float z,
                                       int fma) {
    float dx = __fsub_rn(ux, x), dy = __fsub_rn(uy, y), dz = __fsub_rn(uz, z);
    if (fma) return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// curand_init(seed, 0, 0) followed by one curand_uniform (curand_kernel.h:772-798, 863-874).
__device__ __forceinline__ float xorwow_first_uniform(unsigned long long seed) {
    unsigned s0 = (unsigned)seed ^ 0xaad26b49u;
    unsigned s1 = (unsigned)(seed >> 32) ^ 0xf7dcefddu;
    unsigned t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    unsigned d = 6615241u + t1 + t0;
    unsigned v0 = 123456789u + t0, v4 = 5783321u + t0;
    unsigned t = v0 ^ (v0 >> 2);
    v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
    d += 362437u;
    return __fmaf_rn((float)(v4 + d), 2.3283064e-10f, 2.3283064e-10f / 2.0f);
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Block-wide exclusive scan over n items.  load(i) gives item i, store(i, excl) receives its
// exclusive prefix.  Every thread of the block must call it; `scratch` needs 33 ints of shared
// memory.  Returns the total.  Threads own contiguous chunks, so the result is order-exact.
template <int THREADS, class Load, class Store>
__device__ int block_excl_scan(int n, int *scratch, Load load, Store store) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + THREADS - 1) / THREADS;
    const int beg = min(tid * per, n), end = min(beg + per, n);
    int sum = 0;
    for (int i = beg; i < end; i++) sum += load(i);
    int incl = warp_incl_scan(sum, lane);
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < THREADS / 32 ? scratch[lane] : 0;
        int wi = warp_incl_scan(w, lane);
        scratch[lane] = wi - w;
        if (lane == 31) scratch[32] = wi;
    }
    __syncthreads();
    int run = scratch[warp] + incl - sum;
    for (int i = beg; i < end; i++) {
        int v = load(i);
        store(i, run);
        run += v;
    }
    int total = scratch[32];
    __syncthreads();
    return total;
}

}  // namespace gg
