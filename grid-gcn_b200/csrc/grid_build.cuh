// Voxel-hash build for Gridify / GridifyKNN / GridifyUp (sm_100a).
//
// Replaces reference kernels K1/K3/K5 (gridifyop/gridify.cu:102-191, gridifyknn.cu:115-204,
// gridify_up.cu:102-170) -- NOT a port: the reference claims bucket slots and centre ids with
// global atomics into dense G x P tables (16.4 MB of scratch per 40^3 cloud), which is what makes
// it non-deterministic.  Here one CTA owns one cloud and builds, entirely in shared memory when
// the cloud fits, a compact sorted voxel table (CSR):
//     occupancy bitmap -> popcount prefix (voxel -> dense voxel number)
//     -> per-voxel counts -> exclusive scan -> unordered scatter -> rank sort inside each segment
// so that `sorted` lists every voxel's points in ascending id order (= the canonical schedule's
// bucket order, SURVEY.md s8c rule 1), centres are numbered by first occurrence through a second
// bitmap + popcount prefix over point ids, and the barycentre of each centre voxel is accumulated
// sequentially in fp32 in ascending point order (rule 5).  The scratch written back for the query
// kernel is ~N*4 + O*24 + G/4 bytes per cloud instead of G*P*4.
#pragma once
#include "common.cuh"

namespace gg {

constexpr int kBuildThreads = 1024;
// Segments (points of one voxel) longer than this are ranked by an ordered block scan over the cloud's points instead
// of the per-point O(len) count: a degenerate cloud (every point in one voxel) would otherwise cost O(N^2).
constexpr int kHeavySeg = 1024;
constexpr int kMaxHeavy = 128;  // heavy voxels handled per cloud (N / kHeavySeg <= 128 up to N = 131072); more: quadratic path

// Ranks the points of the heavy voxels of one cloud (called by all threads of a CTA).  key[i] = linear voxel of point i,
// vend[c] = END offset of dense voxel c, tmp = unordered scatter; writes sorted[] and the first-occurrence bits.
template <int THREADS, typename FirstFn>
__device__ __forceinline__ void rank_heavy_segments(const int *key, const int *vend, const int *tmp, int *sorted, int nocc,
                                                    int npts, int *scratch, int *heavy, int *nheavy, FirstFn mark_first) {
    if (threadIdx.x == 0) *nheavy = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < nocc; c += THREADS) {
        const int s = c ? vend[c - 1] : 0;
        if (vend[c] - s > kHeavySeg) {
            const int k = atomicAdd(nheavy, 1);
            if (k < kMaxHeavy) heavy[k] = c;
        }
    }
    __syncthreads();
    const int nh = *nheavy;
    if (nh > kMaxHeavy) return;  // (the per-point pass ranks everything itself in that case)
    for (int h = 0; h < nh; h++) {
        const int c = heavy[h];
        const int s = c ? vend[c - 1] : 0;
        const int lin = key[tmp[s]];
        block_excl_scan<THREADS>(
            npts, scratch, [&](int i) { return key[i] == lin ? 1 : 0; },
            [&](int i, int v) {
                if (key[i] == lin) {
                    sorted[s + v] = i;
                    if (v == 0) mark_first(i);
                }
            });
    }
}

// Shared-memory budget of the build kernel in 32-bit words: bitmap + wordpfx + first-point bitmap
// and its prefix + scan scratch, and (when the cloud fits) key/vend/tmp per point.
__host__ inline size_t build_smem_words(int N, int G, bool points_in_smem) {
    size_t W = (G + 31) / 32, NW = (N + 31) / 32, V = N < G ? N : G;
    size_t words = 2 * W + 2 * NW + 64;
    if (points_in_smem) words += (size_t)N * 2 + V;
    return words;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
grid_build_kernel(const float4 *__restrict__ data, const int *__restrict__ npts_arr, GridParams g,
                  int *__restrict__ ws_base, WsLayout L, float *__restrict__ centmsk,
                  int *__restrict__ centnum, int points_in_smem, int want_centers) {
    extern __shared__ int smem[];
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const int N = g.N, W = g.W, O = g.O;
    const int NW = (N + 31) / 32;
    const int V = N < g.G ? N : g.G;
    int *ws = ws_base + (size_t)b * L.stride;

    unsigned *bitmap = reinterpret_cast<unsigned *>(smem);
    int *wordpfx = smem + W;
    unsigned *firstmap = reinterpret_cast<unsigned *>(smem + 2 * W);
    int *firstpfx = smem + 2 * W + NW;
    int *scratch = smem + 2 * W + 2 * NW;  // 64 words
    int *key, *vend, *tmp;
    if (points_in_smem) {
        key = scratch + 64;
        vend = key + N;
        tmp = vend + V;
    } else {
        key = ws + L.key;
        vend = ws + L.vend;
        tmp = ws + L.tmp;
    }
    int *sorted = ws + L.sorted;
    const float4 *pts = data + (size_t)b * N;
    int npts = npts_arr[b];
    npts = npts < 0 ? 0 : (npts > N ? N : npts);

    // P0: clear bitmaps
    for (int i = tid; i < W; i += THREADS) bitmap[i] = 0u;
    for (int i = tid; i < NW; i += THREADS) firstmap[i] = 0u;
    __syncthreads();

    // P1: voxelise every point (coalesced 128-bit loads), set occupancy bits
    for (int i = tid; i < npts; i += THREADS) {
        float4 p = __ldg(pts + i);
        int lin = voxel_of(p.x, p.y, p.z, g);
        key[i] = lin;
        if (lin >= 0) atomicOr(&bitmap[lin >> 5], 1u << (lin & 31));
    }
    __syncthreads();

    // P2: popcount prefix over bitmap words -> dense voxel numbering in ascending linear index
    const int nocc = block_excl_scan<THREADS>(
        W, scratch, [&](int i) { return __popc(bitmap[i]); }, [&](int i, int v) { wordpfx[i] = v; });

    auto dense_of = [&](int lin) {
        unsigned wbits = bitmap[lin >> 5];
        return wordpfx[lin >> 5] + __popc(wbits & ((1u << (lin & 31)) - 1u));
    };

    // P3: per-voxel point counts (order-independent)
    for (int i = tid; i < nocc; i += THREADS) vend[i] = 0;
    __syncthreads();
    for (int i = tid; i < npts; i += THREADS) {
        int lin = key[i];
        if (lin >= 0) atomicAdd(&vend[dense_of(lin)], 1);
    }
    __syncthreads();

    // P4: counts -> segment starts (cursor)
    block_excl_scan<THREADS>(
        nocc, scratch, [&](int i) { return vend[i]; }, [&](int i, int v) { vend[i] = v; });

    // P5: unordered scatter; afterwards vend[c] is the END offset of segment c
    for (int i = tid; i < npts; i += THREADS) {
        int lin = key[i];
        if (lin >= 0) {
            int pos = atomicAdd(&vend[dense_of(lin)], 1);
            tmp[pos] = i;
        }
    }
    __syncthreads();

    // P6: rank sort inside each segment -> ascending ids; the segment minimum marks the voxel's
    // first occurrence in point order.  Long segments first, by ordered scans (rank_heavy_segments).
    __shared__ int heavy[kMaxHeavy];
    __shared__ int nheavy;
    rank_heavy_segments<THREADS>(key, vend, tmp, sorted, nocc, npts, scratch, heavy, &nheavy, [&](int i) {
        if (want_centers) atomicOr(&firstmap[i >> 5], 1u << (i & 31));
    });
    const bool heavy_done = nheavy <= kMaxHeavy;
    for (int i = tid; i < npts; i += THREADS) {
        int lin = key[i];
        if (lin >= 0) {
            int c = dense_of(lin);
            int s = c ? vend[c - 1] : 0, e = vend[c];
            if (heavy_done && e - s > kHeavySeg) continue;
            int r = 0;
            for (int j = s; j < e; j++) r += (tmp[j] < i);
            sorted[s + r] = i;
            if (r == 0 && want_centers) atomicOr(&firstmap[i >> 5], 1u << (i & 31));
        }
    }
    // publish the lookup tables for the query kernel
    for (int i = tid; i < W; i += THREADS) {
        ws[L.bitmap + i] = (int)bitmap[i];
        ws[L.wordpfx + i] = wordpfx[i];
    }
    if (points_in_smem)
        for (int i = tid; i < nocc; i += THREADS) ws[L.vend + i] = vend[i];
    if (tid == 0) ws[0] = nocc;
    __syncthreads();
    if (!want_centers) return;

    // P7: centre id = rank of the voxel's first point among all first points
    block_excl_scan<THREADS>(
        NW, scratch, [&](int i) { return __popc(firstmap[i]); },
        [&](int i, int v) { firstpfx[i] = v; });

    // P8: per occupied voxel: centre id, linear index, sequential fp32 barycentre sums
    const int ncent = nocc < O ? nocc : O;
    float4 *cent_acc = reinterpret_cast<float4 *>(ws + L.cent_acc);
    for (int c = tid; c < nocc; c += THREADS) {
        int s = c ? vend[c - 1] : 0, e = vend[c];
        int first = sorted[s];
        int cid = firstpfx[first >> 5] + __popc(firstmap[first >> 5] & ((1u << (first & 31)) - 1u));
        if (cid >= O) continue;  // canonical keep-first on centre overflow (gridify.cu:180-187)
        ws[L.cent_lin + cid] = key[first];
        float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
        if (g.loc == 1) {  // gridify.cu:155-162: product rounded, then added, ascending ids
            for (int j = s; j < e; j++) {
                float4 p = __ldg(pts + sorted[j]);
                ax = __fadd_rn(ax, __fmul_rn(p.x, p.w));
                ay = __fadd_rn(ay, __fmul_rn(p.y, p.w));
                az = __fadd_rn(az, __fmul_rn(p.z, p.w));
                aw = __fadd_rn(aw, p.w);
            }
        }
        cent_acc[cid] = make_float4(ax, ay, az, aw);
    }
    // P9: centre mask and count (gridify.cu:179, :222-224)
    // (Gridify_occaware passes null: its sampling kernel writes both after choosing the centres)
    if (centmsk != nullptr)
        for (int o = tid; o < O; o += THREADS) centmsk[(size_t)b * O + o] = o < ncent ? 1.0f : 0.0f;
    if (centnum != nullptr && tid == 0) centnum[b] = ncent;
}

}  // namespace gg
