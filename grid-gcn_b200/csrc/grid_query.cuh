// Per-centre neighbour queries over the sorted voxel table (sm_100a): one warp per centre.
//
// Replaces reference kernels K2 (gridify.cu:193-291), K4 (gridifyknn.cu:206-333) and K6
// (gridify_up.cu:172-225), where ONE THREAD walks up to kernel^3 voxels serially with dependent
// loads (and, for K4, insertion-sorts into a 128-entry local-memory array).  Here the 32 lanes of a
// warp look up the neighbour voxels in parallel (bitmap test + popcount prefix + segment bounds),
// gather candidate rows with 128-bit loads, and select/sort through a warp-level bitonic network
// in shared memory; every output row is written once, coalesced, including the rows beyond
// actual_centnum (their init values, gridify-inl.h:117-121), so no fill kernels are needed.
#pragma once
#include "common.cuh"

namespace gg {

constexpr int kQueryWarps = 8;   // warps per CTA
constexpr int kMaxP = 128;       // reference K4 holds best[128] (gridifyknn.cu:257)

struct CloudTable {
    const unsigned *bitmap;
    const int *wordpfx, *vend, *sorted, *cent_lin;
    const float4 *cent_acc;
};

__device__ __forceinline__ CloudTable cloud_table(const int *ws_base, const WsLayout &L, int b) {
    const int *ws = ws_base + (size_t)b * L.stride;
    CloudTable t;
    t.bitmap = reinterpret_cast<const unsigned *>(ws + L.bitmap);
    t.wordpfx = ws + L.wordpfx;
    t.vend = ws + L.vend;
    t.sorted = ws + L.sorted;
    t.cent_lin = ws + L.cent_lin;
    t.cent_acc = reinterpret_cast<const float4 *>(ws + L.cent_acc);
    return t;
}

// Segment [s, e) of voxel (d,h,w) or an empty range when outside the grid / unoccupied.
__device__ __forceinline__ void voxel_segment(const CloudTable &t, const GridParams &g, int d, int h,
                                              int w, int &s, int &e) {
    s = e = 0;
    if (d < 0 || d >= g.grid[2] || h < 0 || h >= g.grid[1] || w < 0 || w >= g.grid[0]) return;
    int lin = d * (g.grid[0] * g.grid[1]) + h * g.grid[0] + w;
    unsigned bits = t.bitmap[lin >> 5];
    if (!((bits >> (lin & 31)) & 1u)) return;
    int c = t.wordpfx[lin >> 5] + __popc(bits & ((1u << (lin & 31)) - 1u));
    s = c ? t.vend[c - 1] : 0;
    e = t.vend[c];
}

// Sum of (float)(int)w over the first n ids of `row` in slot order (gridify.cu:255-258: the
// weight is read into an `int`).  The terms are integers, so while sum|term| < 2^24 every partial
// fp32 sum is exact and the order is immaterial: reduce exactly in int64.  Otherwise lane 0 redoes
// the reference's sequential fp32 accumulation.
__device__ __forceinline__ float weight_sum(const float4 *pts, const int *row, int n, int lane) {
    long long s = 0, a = 0;
    for (int i = lane; i < n; i += 32) {
        int iw = (int)__ldg(&pts[row[i]].w);
        s += iw;
        a += iw < 0 ? -(long long)iw : iw;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s += __shfl_xor_sync(kFull, s, d);
        a += __shfl_xor_sync(kFull, a, d);
    }
    if (a < (1LL << 24)) return (float)s;
    float f = 0.f;
    if (lane == 0)
        for (int i = 0; i < n; i++) f = __fadd_rn(f, (float)(int)__ldg(&pts[row[i]].w));
    return __shfl_sync(kFull, f, 0);
}

__device__ __forceinline__ float4 center_row(const CloudTable &t, int o, int loc, float wsum) {
    float4 c = make_float4(1.f, 1.f, 1.f, wsum);  // loc==0: xyz keep the init value 1.0
    if (loc == 1) {                                // gridify.cu:280-288
        float4 a = t.cent_acc[o];
        c.x = __fdiv_rn(a.x, a.w);
        c.y = __fdiv_rn(a.y, a.w);
        c.z = __fdiv_rn(a.z, a.w);
    }
    return c;
}

__device__ __forceinline__ void decode_lin(int lin, const GridParams &g, int &c2, int &c1, int &c0) {
    int gxy = g.grid[0] * g.grid[1];
    c2 = lin / gxy;
    c1 = (lin - c2 * gxy) / g.grid[0];
    c0 = lin - c2 * gxy - c1 * g.grid[0];
}

// ------------------------------------------------------------------------------------------------
// Gridify query (A.3): first min(P, total) ids in raster order d -> h -> w, keep-first beyond P.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kQueryWarps * 32)
gridify_query_kernel(const float4 *__restrict__ data, GridParams g, const int *__restrict__ ws_base,
                     WsLayout L, const int *__restrict__ centnum, int *__restrict__ nebidx,
                     float *__restrict__ nebmsk, float4 *__restrict__ cent) {
    __shared__ int rows[kQueryWarps][kMaxP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *row = rows[warp];
    const int P = g.P, O = g.O, ks = g.ks, S = ks * ks * ks, r = (ks - 1) / 2;
    const long long total_centers = (long long)g.B * O;
    for (long long ci = (long long)blockIdx.x * kQueryWarps + warp; ci < total_centers;
         ci += (long long)gridDim.x * kQueryWarps) {
        const int b = (int)(ci / O), o = (int)(ci % O);
        int *out_idx = nebidx + ci * P;
        float *out_msk = nebmsk + ci * P;
        if (o >= centnum[b]) {  // rows beyond actual_centnum keep the init values
            for (int s = lane; s < P; s += 32) {
                out_idx[s] = 0;
                out_msk[s] = 0.f;
            }
            if (lane == 0) cent[ci] = make_float4(1.f, 1.f, 1.f, 1.f);
            continue;
        }
        const CloudTable t = cloud_table(ws_base, L, b);
        const float4 *pts = data + (size_t)b * g.N;
        int c2, c1, c0;
        decode_lin(t.cent_lin[o], g, c2, c1, c0);
        int filled = 0;
        for (int t0 = 0; t0 < S && filled < P; t0 += 32) {
            int tt = t0 + lane, s = 0, e = 0;
            if (tt < S)
                voxel_segment(t, g, tt / (ks * ks) - r + c2, (tt % (ks * ks)) / ks - r + c1,
                              tt % ks - r + c0, s, e);
            int amount = min(P, e - s);  // gridify.cu:249
            int incl = warp_incl_scan(amount, lane);
            int slot = filled + incl - amount;
            for (int j = 0; j < amount && slot + j < P; j++) row[slot + j] = t.sorted[s + j];
            filled += __shfl_sync(kFull, incl, 31);
        }
        __syncwarp();
        const int n = min(filled, P);
        const int pad = row[0];  // initID (gridify.cu:253, :275-279); n >= 1 always
        for (int s = lane; s < P; s += 32) {
            out_idx[s] = s < n ? row[s] : pad;
            out_msk[s] = s < n ? 1.f : 0.f;
        }
        float wsum = weight_sum(pts, row, n, lane);
        if (lane == 0) cent[ci] = center_row(t, o, g.loc, wsum);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-level top-P collector: a buffer of CAP 64-bit keys per warp in shared memory; the first `fill`
// entries are live.  flush() bitonic-sorts them ascending and keeps the first P.  A key is
//     [ sort value : 32 | arrival order of the voxel : vbits | point id : 32 - vbits ]
// so it is unique and self-contained (the id is read back from its low bits: one 8-byte exchange per
// compare instead of key + payload).  Inside a voxel ids ascend in arrival order, so
// (value, voxel order, id) orders exactly like (value, arrival sequence number): the stable order the
// reference's strict-< insertion produces (gridifyknn.cu:288-298).
// ------------------------------------------------------------------------------------------------
template <int CAP>
struct TopP {
    unsigned long long *keys;
    int fill;
    unsigned idmask;  // (1 << idbits) - 1

    __device__ __forceinline__ int id_of(int i) const { return (int)((unsigned)keys[i] & idmask); }

    __device__ __forceinline__ void sort(int lane) {
        int n = 2;
        while (n < fill) n <<= 1;
        for (int i = fill + lane; i < n; i += 32) keys[i] = ~0ull;
        __syncwarp();
        for (int k = 2; k <= n; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int p = lane; p < (n >> 1); p += 32) {
                    const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                    const int q = i | j;
                    const unsigned long long a = keys[i], c = keys[q];
                    const bool up = (i & k) == 0;
                    if ((a > c) == up) {
                        keys[i] = c;
                        keys[q] = a;
                    }
                }
                __syncwarp();
            }
        }
    }
    __device__ __forceinline__ void flush(int P, int lane) {
        if (fill > 1) sort(lane);
        fill = min(fill, P);
    }
};

// Append the candidates of up to 32 voxels (one per lane: segment start `s`, `amount` ids, arrival order
// `vorder` of the voxel) in lane order.  `make_key(id, vorder)` builds the 64-bit key.
template <int CAP, class KeyFn>
__device__ __forceinline__ void append_batch(TopP<CAP> &tp, const int *sorted, int s, int amount,
                                             int P, int lane, int vorder, int &seen, KeyFn make_key) {
    int incl = warp_incl_scan(amount, lane);
    int pfx = incl - amount;
    const int batch_total = __shfl_sync(kFull, incl, 31);
    int base = 0;
    bool done = amount == 0;
    while (__any_sync(kFull, !done)) {
        int room = CAP - tp.fill;
        bool can = !done && (pfx - base + amount <= room);
        if (!__any_sync(kFull, can)) {  // CAP >= 2P guarantees progress after a flush
            tp.flush(P, lane);
            continue;
        }
        if (can) {
            int at = tp.fill + pfx - base;
            for (int j = 0; j < amount; j++) tp.keys[at + j] = make_key(sorted[s + j], vorder);
            done = true;
        }
        int mine = can ? amount : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(kFull, mine, d);
        tp.fill += mine;
        base += mine;
        __syncwarp();
    }
    seen += batch_total;
}

// ------------------------------------------------------------------------------------------------
// GridifyKNN query (A.4): Chebyshev shells around the centre voxel, candidates = first min(P,cnt)
// ids of every voxel in loop order w -> h -> d, stable sort by squared distance to the voxel
// centre in the shifted frame, stop after the first shell with cumulative candidates >= P.
// ------------------------------------------------------------------------------------------------
template <int CAP>
__global__ void __launch_bounds__(kQueryWarps * 32, 3)
gridify_knn_query_kernel(const float4 *__restrict__ data, GridParams g,
                         const int *__restrict__ ws_base, WsLayout L,
                         const int *__restrict__ centnum, int *__restrict__ nebidx,
                         float *__restrict__ nebmsk, float4 *__restrict__ cent) {
    __shared__ unsigned long long s_keys[kQueryWarps][CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = g.P, O = g.O, ks = g.ks;
    // key layout: the voxel arrival order needs vbits, the rest of the low word holds the point id
    int total_combos = 0;
    for (int l = 0; l < (ks + 1) / 2; l++) total_combos += (2 * l + 1) * (2 * l + 1) * (2 * l + 1);
    int vbits = 1;
    while ((1 << vbits) < total_combos) vbits++;
    const int idbits = 32 - vbits;
    const int fma = (g.flags & GRIDGCN_FLAG_DIST_FMA) ? 1 : 0;
    const long long total_centers = (long long)g.B * O;
    for (long long ci = (long long)blockIdx.x * kQueryWarps + warp; ci < total_centers;
         ci += (long long)gridDim.x * kQueryWarps) {
        const int b = (int)(ci / O), o = (int)(ci % O);
        int *out_idx = nebidx + ci * P;
        float *out_msk = nebmsk + ci * P;
        if (o >= centnum[b]) {
            for (int s = lane; s < P; s += 32) {
                out_idx[s] = 0;
                out_msk[s] = 0.f;
            }
            if (lane == 0) cent[ci] = make_float4(1.f, 1.f, 1.f, 1.f);
            continue;
        }
        const CloudTable t = cloud_table(ws_base, L, b);
        const float4 *pts = data + (size_t)b * g.N;
        int c2, c1, c0;
        decode_lin(t.cent_lin[o], g, c2, c1, c0);
        // gridifyknn.cu:253-255: (int + 0.5) * voxel evaluated in double, rounded to float
        const float ux = (float)(((double)c0 + 0.5) * (double)g.voxel[0]);
        const float uy = (float)(((double)c1 + 0.5) * (double)g.voxel[1]);
        const float uz = (float)(((double)c2 + 0.5) * (double)g.voxel[2]);
        TopP<CAP> tp{s_keys[warp], 0, (1u << idbits) - 1u};
        int seq = 0, vbase = 0;
        auto key_of = [&](int id, int vorder) {
            float4 q = __ldg(pts + id);
            float dst = dist2(ux, uy, uz, q.x, q.y, q.z, fma);
            return ((unsigned long long)__float_as_uint(dst) << 32) | ((unsigned)vorder << idbits) | (unsigned)id;
        };
        for (int layer = 0; layer < (ks + 1) / 2; layer++) {
            const int n1 = 2 * layer + 1, combos = n1 * n1 * n1;
            for (int t0 = 0; t0 < combos; t0 += 32) {
                int tt = t0 + lane, s = 0, e = 0;
                if (tt < combos) {
                    int w = tt / (n1 * n1) - layer, h = (tt / n1) % n1 - layer, d = tt % n1 - layer;
                    if (max(max(abs(w), abs(h)), abs(d)) == layer)
                        voxel_segment(t, g, d + c2, h + c1, w + c0, s, e);
                }
                int amount = min(P, e - s);
                append_batch<CAP>(tp, t.sorted, s, amount, P, lane, vbase + tt, seq, key_of);
            }
            vbase += combos;
            // gridifyknn.cu:304-305: need_P -= amount_layer; stop once the cumulative number of
            // candidates (== seq) reaches P
            if (seq >= P) break;
        }
        tp.flush(P, lane);
        __syncwarp();
        const int found = tp.fill;  // >= 1: the centre voxel is never empty
        // unpack the ids in place (low word of every key) so that weight_sum can index them
        int *ids = reinterpret_cast<int *>(tp.keys);
        {
            int v0 = lane < found ? tp.id_of(lane) : 0, v1 = lane + 32 < found ? tp.id_of(lane + 32) : 0;
            int v2 = lane + 64 < found ? tp.id_of(lane + 64) : 0, v3 = lane + 96 < found ? tp.id_of(lane + 96) : 0;
            __syncwarp();
            if (lane < found) ids[lane] = v0;
            if (lane + 32 < found) ids[lane + 32] = v1;
            if (lane + 64 < found) ids[lane + 64] = v2;
            if (lane + 96 < found) ids[lane + 96] = v3;
            __syncwarp();
        }
        const int pad = ids[0];
        for (int s = lane; s < P; s += 32) {
            out_idx[s] = s < found ? ids[s] : pad;  // :308-310, :317-321
            out_msk[s] = 1.f;                       // :312 mask 1 on every slot
        }
        float wsum = weight_sum(pts, ids, found, lane);
        if (lane == 0) cent[ci] = center_row(t, o, g.loc, wsum);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// GridifyUp query (A.5).  The reference splats every down point into the kernel^3 voxels around
// it (K5) so that the up point reads one bucket (K6).  Equivalent without the splat: the bucket of
// voxel u is the ascending-id merge of the down points of u's in-grid neighbour voxels (each point
// reaches u through exactly one offset), truncated to P.  So: gather the first min(P,len) ids of
// every neighbour segment, keep the P smallest ids (key = id), total = sum of full lengths.
// ------------------------------------------------------------------------------------------------
template <int CAP>
__global__ void __launch_bounds__(kQueryWarps * 32)
gridify_up_query_kernel(const float4 *__restrict__ updata, const int *__restrict__ upnum,
                        GridParams g, const int *__restrict__ ws_base, WsLayout L,
                        int *__restrict__ nebidx, float *__restrict__ nebmsk) {
    __shared__ unsigned long long s_keys[kQueryWarps][CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = g.P, O = g.O, ks = g.ks, S = ks * ks * ks, r = (ks - 1) / 2;
    const long long total_rows = (long long)g.B * O;
    for (long long ci = (long long)blockIdx.x * kQueryWarps + warp; ci < total_rows;
         ci += (long long)gridDim.x * kQueryWarps) {
        const int b = (int)(ci / O), o = (int)(ci % O);
        int *out_idx = nebidx + ci * P;
        float *out_msk = nebmsk + ci * P;
        int lin = -1;
        if (o < upnum[b]) {
            float4 p = __ldg(updata + ci);
            lin = voxel_of(p.x, p.y, p.z, g);
        }
        long long count = 0;
        TopP<CAP> tp{s_keys[warp], 0, 0xffffffffu};  // key = id
        if (lin >= 0) {
            const CloudTable t = cloud_table(ws_base, L, b);
            int c2, c1, c0;
            decode_lin(lin, g, c2, c1, c0);
            int seq = 0;
            auto key_of = [&](int id, int) { return (unsigned long long)(unsigned)id; };
            for (int t0 = 0; t0 < S; t0 += 32) {
                int tt = t0 + lane, s = 0, e = 0;
                if (tt < S)
                    voxel_segment(t, g, tt / (ks * ks) - r + c2, (tt % (ks * ks)) / ks - r + c1,
                                  tt % ks - r + c0, s, e);
                int len = e - s;
                long long lsum = len;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) lsum += __shfl_xor_sync(kFull, lsum, d);
                count += lsum;
                append_batch<CAP>(tp, t.sorted, s, min(P, len), P, lane, 0, seq, key_of);
            }
            tp.flush(P, lane);
            __syncwarp();
        }
        const int n = (int)min((long long)P, count);  // gridify_up.cu:212 j < countlimit
        const int pad = n > 0 ? tp.id_of(0) : 0;      // empty bucket: oracle definition, ids 0
        for (int s = lane; s < P; s += 32) {
            out_idx[s] = s < n ? tp.id_of(s) : pad;
            out_msk[s] = s < n ? 1.f : 0.f;
        }
        __syncwarp();
    }
}

}  // namespace gg
