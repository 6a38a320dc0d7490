// Voxel-hash build with POINT-level parallelism (sm_100a): the same sorted voxel table as grid_build.cuh, built
// by a short sequence of grid-wide kernels (B * N threads) instead of one CTA per cloud.
//
// r01: grid_build_kernel owns one cloud per CTA -- right for hundreds of 8192-point clouds, but the shipped
// ScanNet-81920 configuration runs B = 3 clouds per GPU (segmentation/configs/configs.yaml:163,189): 3 of 148 SMs
// busy, every table in global memory, 1.16 ms of a 1.42 ms step (r02j launch list).  The reference launches
// B * N threads for the same work (gridify.cu:369-385).  Same algorithm, hence the same (deterministic,
// schedule-independent) result as the single-CTA kernel:
//     clear -> voxelise + occupancy bits -> popcount prefix -> counts -> scan -> unordered scatter ->
//     rank sort inside each segment (ascending ids = canonical bucket order) + first-point bits ->
//     first-point prefix -> centres (id = first-occurrence rank, sequential fp32 barycentre), mask, count.
// The per-cloud scans stay one CTA per cloud (<= G/32 words, <= min(N, G) counts: microseconds).
#pragma once
#include "common.cuh"

namespace gg {

constexpr int kBmThreads = 256;

__device__ __forceinline__ int bm_dense_of(const int *ws, const WsLayout &L, int lin) {
    const unsigned wbits = (unsigned)__ldg(ws + L.bitmap + (lin >> 5));
    return __ldg(ws + L.wordpfx + (lin >> 5)) + __popc(wbits & ((1u << (lin & 31)) - 1u));
}
__device__ __forceinline__ int bm_npts(const int *npts_arr, int b, int N) {
    const int n = npts_arr[b];
    return n < 0 ? 0 : (n > N ? N : n);
}

__global__ void __launch_bounds__(kBmThreads) bm_clear_kernel(GridParams g, int *ws_base, WsLayout L) {
    const int NW = (g.N + 31) / 32, per = g.W + NW;
    for (long long i = (long long)blockIdx.x * kBmThreads + threadIdx.x; i < (long long)g.B * per;
         i += (long long)gridDim.x * kBmThreads) {
        const int b = (int)(i / per), j = (int)(i % per);
        int *ws = ws_base + (size_t)b * L.stride;
        if (j < g.W) ws[L.bitmap + j] = 0;
        else ws[L.firstmap + (j - g.W)] = 0;
    }
}

__global__ void __launch_bounds__(kBmThreads)
bm_voxelise_kernel(const float4 *__restrict__ data, const int *__restrict__ npts_arr, GridParams g, int *ws_base, WsLayout L) {
    for (long long i = (long long)blockIdx.x * kBmThreads + threadIdx.x; i < (long long)g.B * g.N;
         i += (long long)gridDim.x * kBmThreads) {
        const int b = (int)(i / g.N), j = (int)(i % g.N);
        if (j >= bm_npts(npts_arr, b, g.N)) continue;
        int *ws = ws_base + (size_t)b * L.stride;
        const float4 p = __ldg(data + i);
        const int lin = voxel_of(p.x, p.y, p.z, g);
        ws[L.key + j] = lin;
        if (lin >= 0) atomicOr(reinterpret_cast<unsigned *>(ws + L.bitmap + (lin >> 5)), 1u << (lin & 31));
    }
}

// one CTA per cloud: popcount prefix of the occupancy words, nocc, zeroed counts
__global__ void __launch_bounds__(1024) bm_prefix_kernel(GridParams g, int *ws_base, WsLayout L) {
    __shared__ int scratch[64];
    int *ws = ws_base + (size_t)blockIdx.x * L.stride;
    const int nocc = block_excl_scan<1024>(
        g.W, scratch, [&](int i) { return __popc((unsigned)ws[L.bitmap + i]); }, [&](int i, int v) { ws[L.wordpfx + i] = v; });
    for (int i = threadIdx.x; i < nocc; i += 1024) ws[L.vend + i] = 0;
    if (threadIdx.x == 0) ws[0] = nocc;
}

__global__ void __launch_bounds__(kBmThreads)
bm_count_kernel(const int *__restrict__ npts_arr, GridParams g, int *ws_base, WsLayout L) {
    for (long long i = (long long)blockIdx.x * kBmThreads + threadIdx.x; i < (long long)g.B * g.N;
         i += (long long)gridDim.x * kBmThreads) {
        const int b = (int)(i / g.N), j = (int)(i % g.N);
        if (j >= bm_npts(npts_arr, b, g.N)) continue;
        int *ws = ws_base + (size_t)b * L.stride;
        const int lin = ws[L.key + j];
        if (lin >= 0) atomicAdd(ws + L.vend + bm_dense_of(ws, L, lin), 1);
    }
}

// one CTA per cloud: counts -> segment starts (cursors)
__global__ void __launch_bounds__(1024) bm_scan_kernel(int *ws_base, WsLayout L) {
    __shared__ int scratch[64];
    __shared__ int nheavy;
    int *ws = ws_base + (size_t)blockIdx.x * L.stride;
    const int nocc = ws[0];
    if (threadIdx.x == 0) nheavy = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < nocc; i += 1024)
        if (ws[L.vend + i] > kHeavySeg) atomicAdd(&nheavy, 1);
    __syncthreads();
    if (threadIdx.x == 0) ws[1] = nheavy;  // header word 1: long segments of this cloud (ranked by scans, see below)
    block_excl_scan<1024>(nocc, scratch, [&](int i) { return ws[L.vend + i]; }, [&](int i, int v) { ws[L.vend + i] = v; });
}

__global__ void __launch_bounds__(kBmThreads)
bm_scatter_kernel(const int *__restrict__ npts_arr, GridParams g, int *ws_base, WsLayout L) {
    for (long long i = (long long)blockIdx.x * kBmThreads + threadIdx.x; i < (long long)g.B * g.N;
         i += (long long)gridDim.x * kBmThreads) {
        const int b = (int)(i / g.N), j = (int)(i % g.N);
        if (j >= bm_npts(npts_arr, b, g.N)) continue;
        int *ws = ws_base + (size_t)b * L.stride;
        const int lin = ws[L.key + j];
        if (lin >= 0) ws[L.tmp + atomicAdd(ws + L.vend + bm_dense_of(ws, L, lin), 1)] = j;  // vend[c] ends up the END of segment c
    }
}

__global__ void __launch_bounds__(kBmThreads)
bm_rank_kernel(const int *__restrict__ npts_arr, GridParams g, int *ws_base, WsLayout L, int want_centers) {
    for (long long i = (long long)blockIdx.x * kBmThreads + threadIdx.x; i < (long long)g.B * g.N;
         i += (long long)gridDim.x * kBmThreads) {
        const int b = (int)(i / g.N), j = (int)(i % g.N);
        if (j >= bm_npts(npts_arr, b, g.N)) continue;
        int *ws = ws_base + (size_t)b * L.stride;
        const int lin = ws[L.key + j];
        if (lin < 0) continue;
        const int c = bm_dense_of(ws, L, lin);
        const int s = c ? ws[L.vend + c - 1] : 0, e = ws[L.vend + c];
        if (e - s > kHeavySeg && ws[1] <= kMaxHeavy) continue;  // long segment: ranked by ordered scans (bm_rank_heavy)
        int r = 0;
        for (int q = s; q < e; q++) r += (ws[L.tmp + q] < j);
        ws[L.sorted + s + r] = j;
        if (r == 0 && want_centers) atomicOr(reinterpret_cast<unsigned *>(ws + L.firstmap + (j >> 5)), 1u << (j & 31));
    }
}

// one CTA per cloud: long segments ranked by ordered scans (grid_build.cuh rank_heavy_segments); header word 1 (written
// by bm_scan_kernel) = number of long segments of the cloud, bm_rank_kernel skips them when the list does not overflow.
// Folded into bm_firstpfx_kernel when centres are wanted (one launch less); on its own otherwise.
__device__ __forceinline__ void bm_rank_heavy(const int *__restrict__ npts_arr, const GridParams &g, int *ws, const WsLayout &L,
                                              int b, int want_centers, int *scratch) {
    __shared__ int heavy[kMaxHeavy];
    __shared__ int nheavy;
    if (ws[1] == 0) return;  // uniform over the CTA
    rank_heavy_segments<1024>(ws + L.key, ws + L.vend, ws + L.tmp, ws + L.sorted, ws[0], bm_npts(npts_arr, b, g.N), scratch, heavy,
                              &nheavy, [&](int i) {
                                  if (want_centers) atomicOr(reinterpret_cast<unsigned *>(ws + L.firstmap + (i >> 5)), 1u << (i & 31));
                              });
    __syncthreads();
}
__global__ void __launch_bounds__(1024)
bm_rank_heavy_kernel(const int *__restrict__ npts_arr, GridParams g, int *ws_base, WsLayout L, int want_centers) {
    __shared__ int scratch[64];
    bm_rank_heavy(npts_arr, g, ws_base + (size_t)blockIdx.x * L.stride, L, blockIdx.x, want_centers, scratch);
}

// one CTA per cloud: first-point prefix; centre mask and count (gridify.cu:179, :222-224)
__global__ void __launch_bounds__(1024)
bm_firstpfx_kernel(const int *__restrict__ npts_arr, GridParams g, int *ws_base, WsLayout L, float *__restrict__ centmsk,
                   int *__restrict__ centnum) {
    __shared__ int scratch[64];
    const int b = blockIdx.x;
    int *ws = ws_base + (size_t)b * L.stride;
    bm_rank_heavy(npts_arr, g, ws, L, b, 1, scratch);  // sets the first-occurrence bits of the long segments
    const int NW = (g.N + 31) / 32;
    block_excl_scan<1024>(
        NW, scratch, [&](int i) { return __popc((unsigned)ws[L.firstmap + i]); }, [&](int i, int v) { ws[L.firstpfx + i] = v; });
    const int nocc = ws[0], ncent = nocc < g.O ? nocc : g.O;
    if (centmsk != nullptr)
        for (int o = threadIdx.x; o < g.O; o += 1024) centmsk[(size_t)b * g.O + o] = o < ncent ? 1.0f : 0.0f;
    if (centnum != nullptr && threadIdx.x == 0) centnum[b] = ncent;
}

// one thread per occupied voxel: centre id, linear index, sequential fp32 barycentre sums
__global__ void __launch_bounds__(kBmThreads)
bm_centers_kernel(const float4 *__restrict__ data, GridParams g, int *ws_base, WsLayout L) {
    const int V = g.N < g.G ? g.N : g.G;
    for (long long i = (long long)blockIdx.x * kBmThreads + threadIdx.x; i < (long long)g.B * V;
         i += (long long)gridDim.x * kBmThreads) {
        const int b = (int)(i / V), c = (int)(i % V);
        int *ws = ws_base + (size_t)b * L.stride;
        if (c >= ws[0]) continue;
        const int s = c ? ws[L.vend + c - 1] : 0, e = ws[L.vend + c];
        const int first = ws[L.sorted + s];
        const int cid = ws[L.firstpfx + (first >> 5)] +
                        __popc((unsigned)ws[L.firstmap + (first >> 5)] & ((1u << (first & 31)) - 1u));
        if (cid >= g.O) continue;  // canonical keep-first on centre overflow (gridify.cu:180-187)
        ws[L.cent_lin + cid] = ws[L.key + first];
        const float4 *pts = data + (size_t)b * g.N;
        float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
        if (g.loc == 1) {  // gridify.cu:155-162: product rounded, then added, ascending ids
            for (int q = s; q < e; q++) {
                const float4 p = __ldg(pts + ws[L.sorted + q]);
                ax = __fadd_rn(ax, __fmul_rn(p.x, p.w));
                ay = __fadd_rn(ay, __fmul_rn(p.y, p.w));
                az = __fadd_rn(az, __fmul_rn(p.z, p.w));
                aw = __fadd_rn(aw, p.w);
            }
        }
        reinterpret_cast<float4 *>(ws + L.cent_acc)[cid] = make_float4(ax, ay, az, aw);
    }
}

}  // namespace gg
