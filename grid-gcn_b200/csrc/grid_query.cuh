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
__device__ __forceinline__ float weight_finish(long long s, long long a, const float4 *pts, const int *row, int n, int lane) {
    // warp totals with two REDUX instructions: clamp the per-lane |.| sums to 2^24 (32 lanes: < 2^30); if
    // the clamped total is below 2^24 no lane was clamped, the total is exact and the signed sum fits
    const unsigned at = __reduce_add_sync(kFull, (unsigned)min(a, 1LL << 24));
    const int st = __reduce_add_sync(kFull, (int)s);
    if (at < (1u << 24)) return (float)st;
    float f = 0.f;
    if (lane == 0)
        for (int i = 0; i < n; i++) f = __fadd_rn(f, (float)(int)__ldg(&pts[row[i]].w));
    return __shfl_sync(kFull, f, 0);
}
__device__ __forceinline__ float weight_sum(const float4 *pts, const int *row, int n, int lane) {
    long long s = 0, a = 0;
    for (int i = lane; i < n; i += 32) {
        int iw = (int)__ldg(&pts[row[i]].w);
        s += iw;
        a += iw < 0 ? -(long long)iw : iw;
    }
    return weight_finish(s, a, pts, row, n, lane);
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
// the same row computed by lanes 0..2 in parallel (one IEEE division sequence per warp instead of three); every lane
// must call, lane 0 ends up with the row
__device__ __forceinline__ float4 center_row_warp(const CloudTable &t, int o, int loc, float wsum, int lane) {
    float4 c = make_float4(1.f, 1.f, 1.f, wsum);
    if (loc == 1) {
        const float *a = reinterpret_cast<const float *>(t.cent_acc + o);
        float q = 1.f;
        if (lane < 3) q = __fdiv_rn(a[lane], a[3]);
        c.x = q;
        c.y = __shfl_sync(kFull, q, 1);
        c.z = __shfl_sync(kFull, q, 2);
    }
    return c;
}

// Division by a loop-invariant divisor: m = ceil(2^32 / d) (d >= 2), umulhi over-estimates by at most one.
struct FastDiv {
    int d;
    unsigned m;
    __device__ __forceinline__ explicit FastDiv(int d_) : d(d_), m(d_ > 1 ? 0xFFFFFFFFu / (unsigned)d_ + 1u : 0u) {}
    __device__ __forceinline__ int div(int n) const {  // 0 <= n < 2^31
        if (d <= 1) return n;
        int q = (int)__umulhi((unsigned)n, m);
        if (q * d > n) q--;
        return q;
    }
};
struct LinDecoder {  // linear voxel index -> (c2, c1, c0), two divisions by grid constants
    FastDiv gxy, gx;
    __device__ __forceinline__ explicit LinDecoder(const GridParams &g) : gxy(g.grid[0] * g.grid[1]), gx(g.grid[0]) {}
    __device__ __forceinline__ void operator()(int lin, int &c2, int &c1, int &c0) const {
        c2 = gxy.div(lin);
        const int rem = lin - c2 * gxy.d;
        c1 = gx.div(rem);
        c0 = rem - c1 * gx.d;
    }
};

// ------------------------------------------------------------------------------------------------
// Gridify query (A.3): first min(P, total) ids in raster order d -> h -> w; beyond P the canonical rule
// keeps the first ones.  GRIDGCN_FLAG_STRICT_RESERVOIR instead reproduces the reference's reservoir over
// the later candidates (gridify.cu:259-270): its seed, index_P * size + grid_pntidx, does not depend on the
// schedule, so this part of the reference IS deterministic.  Candidate c (0-based, c >= P) draws
// insrtidx = ceilf(uniform(XORWOW(seed)) * (c + 1)) - 1 and replaces slot insrtidx if it is < P; the
// reference walks the candidates in order, so a slot ends up with the LAST candidate that drew it: the
// draws are independent of each other, hence one pass records the highest c per slot (shared-memory
// atomicMax) and a second one stores the winners' ids.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kQueryWarps * 32)
gridify_query_kernel(const float4 *__restrict__ data, GridParams g, const int *__restrict__ ws_base,
                     WsLayout L, const int *__restrict__ centnum, int *__restrict__ nebidx,
                     float *__restrict__ nebmsk, float4 *__restrict__ cent) {
    __shared__ int rows[kQueryWarps][kMaxP];
    __shared__ int wins[kQueryWarps][kMaxP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *row = rows[warp], *win = wins[warp];
    const int P = g.P, O = g.O, ks = g.ks, S = ks * ks * ks, r = (ks - 1) / 2;
    const bool strict = (g.flags & GRIDGCN_FLAG_STRICT_RESERVOIR) != 0;
    const long long total_centers = (long long)g.B * O;
    const FastDiv odiv(O);
    const LinDecoder decode(g);
    for (long long ci = (long long)blockIdx.x * kQueryWarps + warp; ci < total_centers;
         ci += (long long)gridDim.x * kQueryWarps) {
        const int b = odiv.div((int)ci), o = (int)ci - b * O;  // B * O < 2^31 (host-checked)
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
        decode(t.cent_lin[o], c2, c1, c0);
        // seed of candidate c: (int)(index_P * size + c + 1), evaluated in 32-bit int, widened (gridify.cu:260)
        const unsigned seed0 = (unsigned)ci * (unsigned)P * (unsigned)S;
        auto draw = [&](int c) {  // slot the reservoir assigns to candidate c >= P (may be >= P: dropped)
            const long long sd = (long long)(int)(seed0 + (unsigned)(c + 1));
            return (int)ceilf(__fmul_rn(xorwow_first_uniform((unsigned long long)sd), (float)(c + 1))) - 1;
        };
        if (strict)
            for (int s = lane; s < P; s += 32) win[s] = -1;
        int filled = 0;
        for (int t0 = 0; t0 < S && (strict || filled < P); t0 += 32) {
            int tt = t0 + lane, s = 0, e = 0;
            if (tt < S)
                voxel_segment(t, g, tt / (ks * ks) - r + c2, (tt % (ks * ks)) / ks - r + c1,
                              tt % ks - r + c0, s, e);
            int amount = min(P, e - s);  // gridify.cu:249
            int incl = warp_incl_scan(amount, lane);
            int slot = filled + incl - amount;
            for (int j = 0; j < amount; j++) {
                const int c = slot + j;
                if (c < P) {
                    row[c] = t.sorted[s + j];
                } else if (strict) {
                    const int ins = draw(c);
                    if (ins < P) atomicMax(&win[ins], c);
                } else {
                    break;
                }
            }
            filled += __shfl_sync(kFull, incl, 31);
        }
        __syncwarp();
        if (strict && filled > P) {  // second pass: the winners' ids
            int seen = 0;
            for (int t0 = 0; t0 < S; t0 += 32) {
                int tt = t0 + lane, s = 0, e = 0;
                if (tt < S)
                    voxel_segment(t, g, tt / (ks * ks) - r + c2, (tt % (ks * ks)) / ks - r + c1,
                                  tt % ks - r + c0, s, e);
                int amount = min(P, e - s);
                int incl = warp_incl_scan(amount, lane);
                int slot = seen + incl - amount;
                for (int j = max(0, P - slot); j < amount; j++) {
                    const int c = slot + j, ins = draw(c);
                    if (ins < P && win[ins] == c) row[ins] = t.sorted[s + j];
                }
                seen += __shfl_sync(kFull, incl, 31);
            }
            __syncwarp();
        }
        const int n = min(filled, P);
        const int pad = row[0];  // initID (gridify.cu:253, :275-279); n >= 1 always
        // NOTE (strict mode): initID is the FIRST candidate, which the reservoir may have replaced in
        // slot 0 -- but rows that overflow have no padding, so `pad` is only read when n < P (no overflow)
        for (int s = lane; s < P; s += 32) {
            out_idx[s] = s < n ? row[s] : pad;
            out_msk[s] = s < n ? 1.f : 0.f;
        }
        // cent.w: the reference updates it as += (int)w_new - w_old per replacement (:266-269), which for the
        // integer-valued weights of the pipeline (1.0 inputs, counts afterwards) equals the sum over the
        // final row
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
// Fully unrolled bitonic networks over shared memory (one warp, 64-bit keys).
template <int N, bool DESC>
__device__ __forceinline__ void bitonic_sort(unsigned long long *keys, int lane) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int p0 = 0; p0 < (N >> 1); p0 += 32) {
                const int p = p0 + lane;
                if ((N >> 1) >= 32 || p < (N >> 1)) {
                    const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                    const int q = i | j;
                    const unsigned long long a = keys[i], c = keys[q];
                    const bool asc = ((i & k) == 0) != DESC;
                    if ((a > c) == asc) {
                        keys[i] = c;
                        keys[q] = a;
                    }
                }
            }
            __syncwarp();
        }
    }
}
// The last log2(N) stages only: sorts a bitonic sequence ascending.
template <int N>
__device__ __forceinline__ void bitonic_merge(unsigned long long *keys, int lane) {
#pragma unroll
    for (int j = N >> 1; j > 0; j >>= 1) {
#pragma unroll
        for (int p0 = 0; p0 < (N >> 1); p0 += 32) {
            const int p = p0 + lane;
            if ((N >> 1) >= 32 || p < (N >> 1)) {
                const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                const int q = i | j;
                const unsigned long long a = keys[i], c = keys[q];
                if (a > c) {
                    keys[i] = c;
                    keys[q] = a;
                }
            }
        }
        __syncwarp();
    }
}
template <bool DESC>
__device__ __forceinline__ void bitonic_sort_n(unsigned long long *keys, int n, int lane) {
    switch (n) {
        case 2: bitonic_sort<2, DESC>(keys, lane); break;
        case 4: bitonic_sort<4, DESC>(keys, lane); break;
        case 8: bitonic_sort<8, DESC>(keys, lane); break;
        case 16: bitonic_sort<16, DESC>(keys, lane); break;
        case 32: bitonic_sort<32, DESC>(keys, lane); break;
        case 64: bitonic_sort<64, DESC>(keys, lane); break;
        case 128: bitonic_sort<128, DESC>(keys, lane); break;
        case 256: bitonic_sort<256, DESC>(keys, lane); break;
        default: break;  // n <= 1
    }
}

template <int CAP>
struct TopP {
    unsigned long long *keys;
    int fill;         // after a flush: keys[0, fill) are the best min(P, seen) keys, ascending
    unsigned idmask;  // (1 << idbits) - 1
    int tail;         // candidates appended since the last flush
    bool merged;      // a flush has happened: the tail grows downwards from keys[CAP - 1]

    __device__ __forceinline__ int id_of(int i) const { return (int)((unsigned)keys[i] & idmask); }
    __device__ __forceinline__ int room() const { return CAP - fill - tail; }
    __device__ __forceinline__ int slot(int f) const { return merged ? CAP - 1 - f : f; }

    // Sort what has been collected and keep the best P.  First flush: plain ascending sort of the
    // (power-of-two padded) buffer.  Later flushes: the kept prefix is already sorted, so only the
    // new tail is sorted (descending, at the top end of the buffer), the gap is padded with +inf --
    // ascending prefix, +inf plateau, descending tail is a bitonic sequence -- and one log2(CAP)
    // merge pass finishes the job.
    __device__ __forceinline__ void flush(int P, int lane) {
        if (!merged) {
            int n = 1;
            while (n < tail) n <<= 1;
            for (int i = tail + lane; i < n; i += 32) keys[i] = ~0ull;
            __syncwarp();
            bitonic_sort_n<false>(keys, n, lane);
            fill = min(tail, P);
        } else if (tail > 0) {
            int m = 1;
            while (m < tail) m <<= 1;
            if (m <= CAP - fill) {
                for (int i = fill + lane; i < CAP - tail; i += 32) keys[i] = ~0ull;
                __syncwarp();
                bitonic_sort_n<true>(keys + (CAP - m), m, lane);
                bitonic_merge<CAP>(keys, lane);
            } else {
                for (int i = fill + lane; i < CAP - tail; i += 32) keys[i] = ~0ull;
                __syncwarp();
                bitonic_sort<CAP, false>(keys, lane);
            }
            fill = min(fill + tail, P);
        }
        tail = 0;
        merged = true;
    }
};

// Append the candidates of up to 32 voxels (one per lane: segment start `s`, `amount` ids, arrival order
// `vorder` of the voxel).  The candidates of the batch are numbered flat, 0 .. total-1, and dealt to the
// lanes round-robin -- every lane does the same amount of work whatever the voxel populations are; a
// candidate finds its voxel by a 5-step binary search over the lanes' inclusive prefix (shuffles).
// When the buffer fills it is flushed (sorted, best P kept) and filling continues.
template <int CAP, class KeyFn>
__device__ __forceinline__ void append_batch(TopP<CAP> &tp, const int *sorted, int s, int amount,
                                             int P, int lane, int vorder, int &seen, KeyFn make_key) {
    const int incl = warp_incl_scan(amount, lane);
    const int total = __shfl_sync(kFull, incl, 31);
    int done = 0;
    while (done < total) {  // warp-uniform
        if (tp.room() == 0) tp.flush(P, lane);  // CAP >= 2P: a flush always frees at least P slots
        const int take = min(tp.room(), total - done);
        for (int f0 = 0; f0 < take; f0 += 32) {
            const int f = done + f0 + lane;
            const bool act = f0 + lane < take;
            int v = 0;
#pragma unroll
            for (int st = 16; st > 0; st >>= 1) {
                const int t = __shfl_sync(kFull, incl, v + st - 1);
                if (t <= f) v += st;
            }
            v = min(v, 31);
            const int v_incl = __shfl_sync(kFull, incl, v), v_amount = __shfl_sync(kFull, amount, v);
            const int v_s = __shfl_sync(kFull, s, v), v_order = __shfl_sync(kFull, vorder, v);
            if (act) {
                const int j = f - (v_incl - v_amount);
                tp.keys[tp.slot(tp.tail + f0 + lane)] = make_key(sorted[v_s + j], v_order);
            }
        }
        tp.tail += take;
        done += take;
        __syncwarp();
    }
    seen += total;
}

// ------------------------------------------------------------------------------------------------
// Fast path of the GridifyKNN query (r02): shells 0 + 1 (the 27 voxels around the centre) hold at most CAP
// candidates and no further shell is needed -- the common case (layer 0 of the seg8192 ladder: ~80 candidates
// for P = 64).  The candidates are numbered in arrival order (centre voxel first, then the 26 others in loop
// order, ascending ids inside a voxel) and sorted as 32-bit keys
//        [ d^2 float bits, low log2(CAP) bits dropped | arrival number ]
// through a bitonic network held in REGISTERS (element e = r * 32 + lane: strides >= 32 are register moves,
// strides < 32 one SHFL + one predicated min/max) -- ~260 instructions instead of ~700 for the 64-bit keys in
// shared memory.  The key orders exactly like (d^2, arrival number) -- the stable order of the reference's
// strict-< insertion (gridifyknn.cu:288-298) -- unless two candidates agree in all the kept bits of d^2; such
// a near tie is detected after the sort (adjacent keys with equal high bits) and the centre is handed to the
// general path, which compares the full 64-bit keys.  Bit-exact by construction.
// ------------------------------------------------------------------------------------------------
template <int RR, int R>
__device__ __forceinline__ void bitonic_regs(unsigned (&v)[R], int lane) {
    constexpr int N = 32 * RR;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {  // partner in the same lane (k >= 64: the direction depends on the register only)
#pragma unroll
                for (int r = 0; r < RR; r++) {
                    const int rp = r ^ (j >> 5);
                    if (rp > r) {
                        const bool asc = ((r * 32) & k) == 0;
                        const unsigned lo = min(v[r], v[rp]), hi = max(v[r], v[rp]);
                        v[r] = asc ? lo : hi;
                        v[rp] = asc ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < RR; r++) {
                    const unsigned o = __shfl_xor_sync(kFull, v[r], j);
                    const int e = r * 32 + lane;
                    const bool keep_min = ((e & j) == 0) == ((e & k) == 0);
                    v[r] = keep_min ? min(v[r], o) : max(v[r], o);
                }
            }
        }
    }
}

// Returns false when the centre needs the general path.  On success buf[0, found) holds the neighbour ids in
// output order.  buf: 2 * CAP ints of this warp.
template <int CAP>
__device__ __forceinline__ bool knn_shell01_fast(const CloudTable &t, const GridParams &g, const float4 *pts, int c2,
                                                 int c1, int c0, float ux, float uy, float uz, int fma, int P, int ks,
                                                 int lane, int *buf, int &found, long long &wsum_s, long long &wsum_a,
                                                 bool &all_kept) {
    constexpr int R = CAP / 32;
    constexpr unsigned SEQM = (unsigned)CAP - 1u;
    // lane l looks up voxel tt: the centre voxel (shell 0, tt = 13) first, then the 26 of shell 1 in loop order
    const int tt = lane == 0 ? 13 : (lane <= 13 ? lane - 1 : lane);
    int s = 0, e = 0;
    if (lane < 27) voxel_segment(t, g, tt % 3 - 1 + c2, (tt / 3) % 3 - 1 + c1, tt / 9 - 1 + c0, s, e);
    int amount = min(P, e - s);
    if (__shfl_sync(kFull, amount, 0) >= P && lane != 0) amount = 0;  // shell 0 alone fills the row (:304-305)
    const int incl = warp_incl_scan(amount, lane);
    const int total = __shfl_sync(kFull, incl, 31);
    if (total > CAP || (total < P && ks > 3)) return false;  // too many candidates / further shells needed
    unsigned key[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        key[r] = 0xFFFFFFFFu;
        if (r * 32 < total) {  // warp-uniform
            const int f = r * 32 + lane;  // candidate f: register r of lane f % 32
            int v = 0;                    // its voxel: the first lane whose inclusive prefix exceeds f
#pragma unroll
            for (int st = 16; st > 0; st >>= 1) {
                const int tpre = __shfl_sync(kFull, incl, v + st - 1);
                if (tpre <= f) v += st;
            }
            v = min(v, 31);
            const int v_incl = __shfl_sync(kFull, incl, v), v_amount = __shfl_sync(kFull, amount, v);
            const int v_s = __shfl_sync(kFull, s, v);
            if (f < total) {
                const int id = t.sorted[v_s + (f - (v_incl - v_amount))];
                const float4 q = __ldg(pts + id);
                const float dst = dist2(ux, uy, uz, q.x, q.y, q.z, fma);
                key[r] = (__float_as_uint(dst) & ~SEQM) | (unsigned)f;
                buf[CAP + f] = id;
                const int iw = (int)q.w;  // weight of the candidate: the whole row's sum when every candidate is kept
                wsum_s += iw;
                wsum_a += iw < 0 ? -(long long)iw : iw;
            }
        }
    }
    if (total <= 32) bitonic_regs<1>(key, lane);
    else if (R >= 2 && total <= 64) bitonic_regs<(R >= 2 ? 2 : 1)>(key, lane);
    else if (R >= 4 && total <= 128) bitonic_regs<(R >= 4 ? 4 : 1)>(key, lane);
    else bitonic_regs<R>(key, lane);
    // near ties: adjacent keys that agree in every kept bit of d^2
    bool tie = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const unsigned nxt_reg = __shfl_sync(kFull, r + 1 < R ? key[r + 1 < R ? r + 1 : r] : 0xFFFFFFFFu, 0);
        unsigned nk = __shfl_down_sync(kFull, key[r], 1);
        if (lane == 31) nk = nxt_reg;
        if (r * 32 + lane + 1 < total && ((key[r] ^ nk) & ~SEQM) == 0u) tie = true;
    }
    if (__any_sync(kFull, tie)) return false;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; r++)
        if (r * 32 + lane < total) buf[r * 32 + lane] = buf[CAP + (key[r] & SEQM)];
    __syncwarp();
    found = min(total, P);
    all_kept = total <= P;
    return true;
}

// ------------------------------------------------------------------------------------------------
// GridifyKNN query (A.4): Chebyshev shells around the centre voxel, candidates = first min(P,cnt)
// ids of every voxel in loop order w -> h -> d, stable sort by squared distance to the voxel
// centre in the shifted frame, stop after the first shell with cumulative candidates >= P.
// ------------------------------------------------------------------------------------------------
#ifndef GG_KNN_MIN_CTAS
#define GG_KNN_MIN_CTAS 4  // resident CTAs per SM the register allocation is bounded for (tools/build_variant.py)
#endif
template <int CAP>
__global__ void __launch_bounds__(kQueryWarps * 32, GG_KNN_MIN_CTAS)
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
    const FastDiv odiv(O);
    const LinDecoder decode(g);
    for (long long ci = (long long)blockIdx.x * kQueryWarps + warp; ci < total_centers;
         ci += (long long)gridDim.x * kQueryWarps) {
        const int b = odiv.div((int)ci), o = (int)ci - b * O;  // B * O < 2^31 (host-checked)
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
        decode(t.cent_lin[o], c2, c1, c0);
        // gridifyknn.cu:253-255: (int + 0.5) * voxel evaluated in double, rounded to float
        const float ux = (float)(((double)c0 + 0.5) * (double)g.voxel[0]);
        const float uy = (float)(((double)c1 + 0.5) * (double)g.voxel[1]);
        const float uz = (float)(((double)c2 + 0.5) * (double)g.voxel[2]);
        int found = 0;
        long long wsum_s = 0, wsum_a = 0;  // per-lane weight sums over ALL candidates of the fast path
        bool all_kept = false;             // ... which are the row's weights when no candidate was dropped
        int *ids = reinterpret_cast<int *>(s_keys[warp]);  // CAP 64-bit keys == 2 * CAP ints
        if (!(ks >= 3 && knn_shell01_fast<CAP>(t, g, pts, c2, c1, c0, ux, uy, uz, fma, P, ks, lane, ids, found, wsum_s,
                                               wsum_a, all_kept))) {
        all_kept = false;
        __syncwarp();
        TopP<CAP> tp{s_keys[warp], 0, (1u << idbits) - 1u, 0, false};
        int seq = 0, vbase = 0;
        auto key_of = [&](int id, int vorder) {
            float4 q = __ldg(pts + id);
            float dst = dist2(ux, uy, uz, q.x, q.y, q.z, fma);
            return ((unsigned long long)__float_as_uint(dst) << 32) | ((unsigned)vorder << idbits) | (unsigned)id;
        };
        int first_layer = 0;
        if (ks >= 3) {
            // shells 0 and 1 in one pass over the 27 voxels around the centre (constant divisors, one
            // table lookup per lane).  The centre voxel is shell 0 (arrival order 0); the other 26 are
            // shell 1 in loop order (arrival order 1 + tt) and only count when shell 0 alone holds
            // fewer than P candidates (gridifyknn.cu:304-305).
            int s = 0, e = 0;
            if (lane < 27)
                voxel_segment(t, g, lane % 3 - 1 + c2, (lane / 3) % 3 - 1 + c1, lane / 9 - 1 + c0, s, e);
            int amount = min(P, e - s);
            if (__shfl_sync(kFull, amount, 13) >= P && lane != 13) amount = 0;
            append_batch<CAP>(tp, t.sorted, s, amount, P, lane, lane == 13 ? 0 : 1 + lane, seq, key_of);
            vbase = 28;
            first_layer = 2;
        }
        for (int layer = first_layer; layer < (ks + 1) / 2 && seq < P; layer++) {
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
            // candidates (== seq) reaches P (the loop condition)
        }
        tp.flush(P, lane);
        __syncwarp();
        found = tp.fill;  // >= 1: the centre voxel is never empty
        // unpack the ids in place (low word of every key) so that weight_sum can index them
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
        }  // general path
        const int pad = ids[0];
        for (int s = lane; s < P; s += 32) {
            out_idx[s] = s < found ? ids[s] : pad;  // :308-310, :317-321
            out_msk[s] = 1.f;                       // :312 mask 1 on every slot
        }
        const float wsum = all_kept ? weight_finish(wsum_s, wsum_a, pts, ids, found, lane) : weight_sum(pts, ids, found, lane);
        const float4 crow = center_row_warp(t, o, g.loc, wsum, lane);
        if (lane == 0) cent[ci] = crow;
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
    const FastDiv odiv(O);
    const LinDecoder decode(g);
    for (long long ci = (long long)blockIdx.x * kQueryWarps + warp; ci < total_rows;
         ci += (long long)gridDim.x * kQueryWarps) {
        const int b = odiv.div((int)ci), o = (int)ci - b * O;  // B * O < 2^31 (host-checked)
        int *out_idx = nebidx + ci * P;
        float *out_msk = nebmsk + ci * P;
        int lin = -1;
        if (o < upnum[b]) {
            float4 p = __ldg(updata + ci);
            lin = voxel_of(p.x, p.y, p.z, g);
        }
        long long count = 0;
        TopP<CAP> tp{s_keys[warp], 0, 0xffffffffu, 0, false};  // key = id
        if (lin >= 0) {
            const CloudTable t = cloud_table(ws_base, L, b);
            int c2, c1, c0;
            decode(lin, c2, c1, c0);
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
