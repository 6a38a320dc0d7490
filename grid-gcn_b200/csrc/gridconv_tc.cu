// Fused GridConv layer on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in
// TMEM, weight slices streamed by the TMA engine's bulk copies) -- GRIDGCN_PRECISION_TF32 / TF32X3.
//
// What is computed (reference segmentation/models/gcn_module_g_att.py:172-287, see
// gridconv_common.cuh) is restructured for the hardware instead of translated:
//
//  * Feature MLP hoisting.  For layers with input features (has_feats, localfdim == 0) the feature
//    MLP of verts_pair_func (:135) is a function of the gathered neighbour ROW only, so
//    MLP(gather(table)) == gather(MLP(table)): kernel A applies it once per source point (B*Nprev
//    rows) instead of once per edge (B*O*K rows, 16x more at K=64), writing a transformed table F.
//    Kernel A has two variants: `point_mlp_rows_tc_kernel` (D[row, ch], epilogue thread = row, used
//    when a [128 x K] hi/lo activation image fits, K <= 128) and `point_mlp_tc_kernel` (D^T[ch, row],
//    64-row tiles, for K = 256).  The per-edge work that remains is what really depends on the
//    (centre, neighbour) pair: the attention MLP, the product and the max pool.
//  * Kernel B (`edge_tc_kernel<NSPLIT, FIRST>`), persistent CTAs, tiles of 128 edges:
//      gather phase (thread = edge; index / row / centre prefetched a tile ahead): neighbour xyz with
//      one 128-bit load -> geo / att_vec in registers -> the tiny-K stages on the CUDA cores in exact
//      fp32 (attention stage 0, K <= 10; first-layer feature stage 0, K = 3: a tensor-core round trip
//      would cost more than the 11 FMAs) -> hi/lo K-major operand images in shared memory;
//      first layer only: the remaining hidden feature stage on the tensor core, D[edge, ch];
//      last attention stage (and first-layer last feature stage) TRANSPOSED, D^T[ch, edge] = W * H^T:
//      the K edges of a centre are K consecutive TMEM columns of ONE lane, so the max pool is a run
//      of FMNMX in one thread's registers -- no shuffles, no shared-memory transpose -- and thread =
//      channel makes the output store coalesced; relu(att) * feat (feat from TMEM for the first layer,
//      gathered from F otherwise), pre-ReLU, centre mask, one [cent | feats] row per centre.
//  * fp32 parity on tf32 tensor cores: TF32X3 splits every operand into hi + lo and issues lo*hi,
//    hi*lo, hi*hi into the same TMEM accumulator (error ~2^-21, fp32-class).  Activations use the
//    measured truncation of the low 13 mantissa bits by the tensor core (hi = raw bits, lo = v -
//    trunc(v): 2 instructions; tc_common.cuh split_op, pinned by tests/test_gpu_tc.py); weights are
//    packed once with round-to-nearest parts.  TF32 issues hi*hi only (speed option, held to 1e-2).
//
// Synchronisation is deliberately lock-step (one __syncthreads per phase, one mbarrier for "MMAs of
// this stage retired", full/empty mbarriers on the weight ring); overlap comes from 2-3 co-resident
// CTAs per SM (TMEM: 256 columns for the first layer, 128 otherwise).  Every mbarrier wait is bounded
// and traps instead of hanging.  `gridgcn_debug_phase_buffer` exposes per-phase cycle counters.
#include "gridconv_common.cuh"
#include "gridconv_tc.cuh"

#include <algorithm>
#include <cstdlib>

namespace gg {

// ------------------------------------------------------------------------------------------------
// Weight packing: folded fp32 W(Cout, Cin) -> hi/lo tf32 operand images in the layouts the MMA
// descriptors expect (tc_common.cuh).  Plain stage: one [Np x Kp] image (LBO = Np*16).  Transposed
// stage: for every 128-row chunk, for every 32-wide k slice, a contiguous [128 x kw] hi image
// followed by its lo image (LBO = 2048), in the order the kernel streams them.
// ------------------------------------------------------------------------------------------------
__global__ void pack_stage_kernel(const float *__restrict__ W, float *__restrict__ dst, TcStage st) {
    const int total = st.Np * st.Kp;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int n = e / st.Kp, k = e % st.Kp;
        const float w = (n < st.Cout && k < st.Cin) ? W[(size_t)n * st.Cin + k] : 0.f;
        float hi, lo;
        tc::split_tf32(w, hi, lo);
        size_t off_hi, off_lo;
        if (!st.transposed) {
            size_t o = (size_t)(k >> 2) * st.Np * 4 + (size_t)(n >> 3) * 32 + (n & 7) * 4 + (k & 3);
            off_hi = o;
            off_lo = (size_t)st.Np * st.Kp + o;
        } else {
            const int j = n >> 7, r = n & 127, t = k / kSliceK, kk = k % kSliceK;
            const int kw = min(kSliceK, st.Kp - t * kSliceK);
            size_t base = (size_t)j * 2 * 128 * st.Kp + (size_t)t * 2 * 128 * kSliceK;
            size_t o = (size_t)(kk >> 2) * 128 * 4 + (size_t)(r >> 3) * 32 + (r & 7) * 4 + (kk & 3);
            off_hi = base + o;
            off_lo = base + (size_t)128 * kw + o;
        }
        dst[st.w_off + off_hi] = hi;
        dst[st.w_off + off_lo] = lo;
    }
}

constexpr int kMaxSeq = 128;  // slices per tile (3 stages x <= 4 chunks x <= 8 slices)

// Weight ring.  All bookkeeping is done by ONE thread (the MMA issuer) and sits on its critical path between
// two MMA batches, so it is kept to 32-bit counters with wrap-around increments (no divisions) and a per-tile
// slice table precomputed in shared memory (r01: the 64-bit modulo / SliceSeq::next version cost ~1.5 us per
// slice, more than the copy and the MMAs together).
struct Ring {
    uint8_t *slots;
    uint32_t off[kMaxRing];  // byte offset of every slot (exact-fit in sticky mode, 32 KB apart otherwise)
    uint64_t *full, *empty;
    int nslots;
    const float *packed;
    const uint32_t *seq_off, *seq_bytes;  // per-tile slice sequence: float offset into packed, bytes
    int per_tile;
    uint32_t remaining;            // producer: slices still to request (whole kernel)
    int p_slot, p_seq;             // producer cursors: ring slot, position in the per-tile sequence
    uint32_t p_round;              // producer: times the slot cursor has wrapped
    int c_slot;                    // consumer cursor
    uint32_t c_round, c_count;     // consumer: wraps; slices consumed, saturating at 2
    __device__ void reset() {
        remaining = 0;
        p_slot = p_seq = c_slot = 0;
        p_round = c_round = c_count = 0;
    }
};

// ---- slice sequence (setup only: the hot path reads the table built from it) ------------------------
struct SliceSeq {
    const TcStage *st[3];
    int n;            // number of transposed stages per tile
    int chunk_major;  // 0: stage by stage (kernel A); 1: chunk by chunk across the stages (kernel B)
    int j, t, s;      // cursor: chunk, slice, stage
    __device__ void reset() { j = t = s = 0; }
    __device__ int per_tile() const {
        int c = 0;
        for (int i = 0; i < n; i++) c += (st[i]->Np / 128) * ((st[i]->Kp + kSliceK - 1) / kSliceK);
        return c;
    }
    // current slice source (float offset into packed) and size in bytes, then advance (wraps per tile)
    __device__ void next(long long &off, uint32_t &bytes) {
        const TcStage &q = *st[s];
        const int kw = min(kSliceK, q.Kp - t * kSliceK);
        off = q.w_off + (long long)j * 2 * 128 * q.Kp + (long long)t * 2 * 128 * kSliceK;
        bytes = (uint32_t)(2 * 128 * kw * 4);
        const int nslice = (q.Kp + kSliceK - 1) / kSliceK;
        if (++t < nslice) return;
        t = 0;
        if (!chunk_major) {
            if (++j == q.Np / 128) {
                j = 0;
                if (++s == n) s = 0;
            }
        } else {
            if (++s == n) {
                s = 0;
                if (++j == q.Np / 128) j = 0;
            }
        }
    }
};

// Producer set-up (the issuing thread only): build the slice table, then prefetch the first slices.
__device__ __forceinline__ void ring_request(Ring &ring, bool wait_empty);
__device__ __forceinline__ void ring_start(Ring &ring, SliceSeq seq, uint32_t *seq_off, uint32_t *seq_bytes,
                                           const float *packed, int per_tile, long long total_slices,
                                           bool sticky) {
    seq.reset();
    for (int i = 0; i < per_tile && i < kMaxSeq; i++) {
        long long off;
        uint32_t bytes;
        seq.next(off, bytes);
        seq_off[i] = (uint32_t)off;
        seq_bytes[i] = bytes;
    }
    ring.packed = packed;
    ring.seq_off = seq_off;
    ring.seq_bytes = seq_bytes;
    ring.per_tile = per_tile;
    const long long pre = min((long long)(sticky ? per_tile : ring.nslots), total_slices);
    ring.remaining = (uint32_t)(sticky ? pre : total_slices);
    for (long long i = 0; i < pre; i++) ring_request(ring, false);
}

// Single-thread producer step: request the next slice of the CTA's sequence into the ring.
__device__ __forceinline__ void ring_request(Ring &ring, bool wait_empty) {
    if (ring.remaining == 0) return;
    const int slot = ring.p_slot;
    if (wait_empty && ring.p_round > 0) wait_bar(&ring.empty[slot], (ring.p_round - 1) & 1u);
    const uint32_t bytes = ring.seq_bytes[ring.p_seq];
    tc::mbar_expect_tx(&ring.full[slot], bytes);
    tc::bulk_g2s(ring.slots + ring.off[slot], ring.packed + ring.seq_off[ring.p_seq], bytes, &ring.full[slot]);
    if (++ring.p_seq == ring.per_tile) ring.p_seq = 0;
    if (++ring.p_slot == ring.nslots) {
        ring.p_slot = 0;
        ring.p_round++;
    }
    ring.remaining--;
}

// Issue every MMA of one transposed stage (single thread).  D^T[128 ch of chunk j, ncols] (+)=
// W_slice * Xop^T; x_hi/x_lo are the shared addresses of the B operand images ([ncols rows x Kp],
// panel stride x_lbo); weight slices arrive through the ring in the packed order.
// SWAP == true: the same slices serve as the B operand and the image as the A operand, i.e.
// D[row, 128 ch of chunk j] (+)= X * W_slice^T (rows in TMEM lanes, channels in columns).
template <int NSPLIT, bool SWAP = false>
__device__ __forceinline__ void run_transposed_stage(int Kp, int nchunk, Ring &ring, bool sticky,
                                                     uint32_t x_hi, uint32_t x_lo, uint32_t x_lbo, int ncols,
                                                     uint32_t tmem_base, int col_stride) {
    const uint32_t idesc = tc::make_idesc_tf32(128, ncols);
    const int nslice = (Kp + kSliceK - 1) / kSliceK;
    for (int j = 0; j < nchunk; j++) {
        const uint32_t d = tmem_base + j * col_stride;
        uint32_t acc = 0;
        for (int t = 0; t < nslice; t++) {
            const int kw = min(kSliceK, Kp - t * kSliceK);
            const int slot = ring.c_slot;
            if (!sticky || ring.c_round == 0) wait_bar(&ring.full[slot], ring.c_round & 1u);
            const uint32_t a_hi = tc::smem_u32(ring.slots + ring.off[slot]);
            const uint32_t a_lo = a_hi + 128 * kw * 4;
            for (int ks = 0; ks < kw / 8; ks++) {
                const uint64_t ah = tc::make_sdesc(a_hi + ks * 2 * 2048, 2048);
                const uint32_t xo = (uint32_t)(t * (kSliceK / 4) + ks * 2) * x_lbo;
                const uint64_t bh = tc::make_sdesc(x_hi + xo, x_lbo);
                if (NSPLIT == 3) {
                    const uint64_t al = tc::make_sdesc(a_lo + ks * 2 * 2048, 2048);
                    const uint64_t bl = tc::make_sdesc(x_lo + xo, x_lbo);
                    if (SWAP) {
                        tc::mma_tf32(d, bl, ah, idesc, acc);
                        tc::mma_tf32(d, bh, al, idesc, 1);
                    } else {
                        tc::mma_tf32(d, al, bh, idesc, acc);
                        tc::mma_tf32(d, ah, bl, idesc, 1);
                    }
                    acc = 1;
                }
                if (SWAP) tc::mma_tf32(d, bh, ah, idesc, acc);
                else tc::mma_tf32(d, ah, bh, idesc, acc);
                acc = 1;
            }
            if (++ring.c_slot == ring.nslots) {
                ring.c_slot = 0;
                ring.c_round++;
            }
            if (ring.c_count < 2) ring.c_count++;
            if (!sticky) {
                tc::mma_commit(&ring.empty[slot]);
                // Keep the ring full, one slice behind: the slot recycled here belongs to the slice
                // BEFORE the one just issued, so waiting for it to drain never idles the tensor pipe.
                if (ring.c_count >= 2) ring_request(ring, true);
            }
        }
    }
}

// Issue the MMAs of one plain stage (single thread): D[128 rows, Np] = X * W^T, W resident.
template <int NSPLIT>
__device__ __forceinline__ void run_plain_stage(const TcStage &st, uint32_t x_hi, uint32_t x_lo,
                                                uint32_t x_lbo, uint32_t w_hi, uint32_t tmem_d) {
    const uint32_t idesc = tc::make_idesc_tf32(128, st.Np);
    const uint32_t w_lbo = (uint32_t)st.Np * 16, w_lo = w_hi + (uint32_t)st.Np * st.Kp * 4;
    uint32_t acc = 0;
    for (int ks = 0; ks < st.Kp / 8; ks++) {
        const uint64_t ah = tc::make_sdesc(x_hi + ks * 2 * x_lbo, x_lbo);
        const uint64_t bh = tc::make_sdesc(w_hi + ks * 2 * w_lbo, w_lbo);
        if (NSPLIT == 3) {
            const uint64_t al = tc::make_sdesc(x_lo + ks * 2 * x_lbo, x_lbo);
            const uint64_t bl = tc::make_sdesc(w_lo + ks * 2 * w_lbo, w_lbo);
            tc::mma_tf32(tmem_d, al, bh, idesc, acc);
            tc::mma_tf32(tmem_d, ah, bl, idesc, 1);
            acc = 1;
        }
        tc::mma_tf32(tmem_d, ah, bh, idesc, acc);
        acc = 1;
    }
}

// TMEM epilogue of a plain stage, thread = row: x = relu(D + bias) -> hi/lo images [128 x Kp_next].
// `bias_s`: the stage's bias in shared memory, zero padded to 128 (global bias loads on this path cost an
// L2 round trip per batch when L1 is thrashed by the gathers).  Columns >= Cout come out as relu(0+0)=0
// because the padded weight rows are zero.
template <int NSPLIT>
__device__ __forceinline__ void plain_epilogue(const TcStage &st, uint32_t tmem_lane_addr, uint32_t row_off,
                                               uint8_t *img_hi, uint8_t *img_lo, uint32_t lbo,
                                               int kp_next, const float *bias_s, int c_begin = 0,
                                               int c_end = 1 << 30) {
    (void)st;
    for (int c0 = c_begin; c0 < min(kp_next, c_end); c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld16(tmem_lane_addr + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = c0 + q * 4;
            if (c < kp_next) {
                const float4 b4 = *reinterpret_cast<const float4 *>(bias_s + c);
                float hi[4], lo[4];
                tc::split_op<NSPLIT>(fmaxf(__uint_as_float(v[q * 4 + 0]) + b4.x, 0.f), hi[0], lo[0]);
                tc::split_op<NSPLIT>(fmaxf(__uint_as_float(v[q * 4 + 1]) + b4.y, 0.f), hi[1], lo[1]);
                tc::split_op<NSPLIT>(fmaxf(__uint_as_float(v[q * 4 + 2]) + b4.z, 0.f), hi[2], lo[2]);
                tc::split_op<NSPLIT>(fmaxf(__uint_as_float(v[q * 4 + 3]) + b4.w, 0.f), hi[3], lo[3]);
                const uint32_t off = row_off + (uint32_t)(c >> 2) * lbo;
                *reinterpret_cast<float4 *>(img_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                if (NSPLIT == 3)
                    *reinterpret_cast<float4 *>(img_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel A: per-point feature MLP, F = MLP_pt(table[:, 4:]) for B*Nprev rows.
// Every stage transposed: D^T[ch, row]; thread = channel in the epilogue.
// ------------------------------------------------------------------------------------------------
template <int NSPLIT>
__global__ void __launch_bounds__(kRowsThreads, 1)
point_mlp_tc_kernel(const __grid_constant__ TcParams p, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[2 * kMaxRing + 1];
    __shared__ uint32_t seq_off_s[kMaxSeq], seq_bytes_s[kMaxSeq];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int TR = p.a_rows;
    int kmax = 0;
    for (int s = 0; s < p.na; s++) kmax = max(kmax, p.a[s].Kp);
    const uint32_t x_lbo = (uint32_t)TR * 16 + 16;  // +16: conflict-free transposed 4-byte stores
    const uint32_t x_img = (uint32_t)(kmax / 4) * x_lbo;
    uint8_t *x_hi = smem, *x_lo = smem + x_img;
    Ring ring;
    ring.slots = smem + pad_to((NSPLIT == 3 ? 2 : 1) * x_img, 128);
    ring.nslots = p.ring_slots;
    ring.full = bars;
    ring.empty = bars + kMaxRing;
    ring.reset();
    for (int i = 0; i < kMaxRing; i++) ring.off[i] = (uint32_t)i * kSlotBytes;
    uint64_t *bar_mma = bars + 2 * kMaxRing;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) {
        for (int i = 0; i < ring.nslots; i++) {
            tc::mbar_init(&ring.full[i], 1);
            tc::mbar_init(&ring.empty[i], 1);
        }
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init_fence();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    SliceSeq prod;
    prod.n = p.na;
    prod.chunk_major = 0;
    for (int s = 0; s < p.na && s < 3; s++) prod.st[s] = &p.a[s];
    prod.reset();
    // NOTE: kernel A supports at most 3 stages in the slice sequence (len(pt_mlp_lst) <= 3)
    const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total_slices = (long long)my_tiles * prod.per_tile();
    const bool sticky = prod.per_tile() <= ring.nslots;  // whole sequence fits: load once, keep
    if (sticky) ring.nslots = max(prod.per_tile(), 1);
    if (warp == 8 && lane == 0) {
        ring_start(ring, prod, seq_off_s, seq_bytes_s, p.packed, prod.per_tile(), total_slices, sticky);
    }
    uint32_t mma_phase = 0;
    const long long rows_total = (long long)p.c.B * p.c.Nprev;
    const int row_w = 4 + p.c.Cin;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long row0 = (long long)tile * TR;
        // ---- load X0 = table[row0 .. row0+TR, 4:4+Cin] as hi/lo K-major images ----
        if (warp < 8) {
            const int kp0 = p.a[0].Kp;
            for (int e = tid; e < TR * (kp0 / 4); e += 256) {
                const int r = e / (kp0 / 4), c = (e % (kp0 / 4)) * 4;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (row0 + r < rows_total) {
                    const float *src = p.c.table + (row0 + r) * row_w + 4 + c;
                    if ((row_w & 3) == 0 && c + 3 < p.c.Cin) {
                        float4 t4 = __ldg(reinterpret_cast<const float4 *>(src));
                        v[0] = t4.x; v[1] = t4.y; v[2] = t4.z; v[3] = t4.w;
                    } else {
                        for (int i = 0; i < 4; i++)
                            if (c + i < p.c.Cin) v[i] = __ldg(src + i);
                    }
                }
                float hi[4], lo[4];
                for (int i = 0; i < 4; i++) tc::split_op<NSPLIT>(v[i], hi[i], lo[i]);
                const uint32_t off = tc::kmajor_off(r, c, x_lbo);
                *reinterpret_cast<float4 *>(x_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                if (NSPLIT == 3) *reinterpret_cast<float4 *>(x_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        for (int s = 0; s < p.na; s++) {
            const TcStage &st = p.a[s];
            if (warp == 8) {
                if (lane == 0) {
                    run_transposed_stage<NSPLIT>(st.Kp, st.Np / 128, ring, sticky, tc::smem_u32(x_hi),
                                                 tc::smem_u32(x_lo), x_lbo, TR, tmem, TR);
                    tc::mma_commit(bar_mma);
                }
                __syncwarp();
            }
            wait_bar(bar_mma, mma_phase);
            mma_phase ^= 1;
            tc::fence_after_sync();
            if (warp < 8) {
                const bool last = s + 1 == p.na;
                const int kp_next = last ? 0 : p.a[s + 1].Kp;
                for (int j = 0; j < st.Np / 128; j++) {
                    const int ch = j * 128 + (tid & 127);
                    const float bias = ch < st.Cout ? __ldg(st.bias + ch) : 0.f;
                    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + j * TR;
                    for (int r0 = (warp >> 2) * (TR / 2); r0 < ((warp >> 2) + 1) * (TR / 2); r0 += 16) {  // two warps per lane quadrant: half of the rows each
                        uint32_t v[16];
                        tc::tmem_ld16(taddr + r0, v);
                        tc::tmem_ld_wait();
                        if (last) {
                            if (ch < st.Cout) {
#pragma unroll
                                for (int i = 0; i < 16; i++)
                                    if (row0 + r0 + i < rows_total)
                                        p.ftab[(row0 + r0 + i) * st.Cout + ch] =
                                            fmaxf(__uint_as_float(v[i]) + bias, 0.f);
                            }
                        } else if (ch < kp_next) {
#pragma unroll
                            for (int i = 0; i < 16; i++) {
                                float x = ch < st.Cout ? fmaxf(__uint_as_float(v[i]) + bias, 0.f) : 0.f;
                                float hi, lo;
                                tc::split_op<NSPLIT>(x, hi, lo);
                                const uint32_t off = tc::kmajor_off(r0 + i, ch, x_lbo);
                                *reinterpret_cast<float *>(x_hi + off) = hi;
                                if (NSPLIT == 3) *reinterpret_cast<float *>(x_lo + off) = lo;
                            }
                        }
                    }
                }
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
            tc::fence_after_sync();
        }
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
// Kernel A, row-major variant (used when the widest stage input fits a [128 x K] hi/lo image, K <= 128):
// D[row, ch] = X * W^T with the rows in the TMEM lanes.  The epilogue thread owns a ROW, so the next
// stage's operand image is written with 16-byte vector stores and every one of the 128 epilogue threads
// works whatever the stage width -- ~4x fewer instructions than the transposed kernel above, which
// remains for the layers whose activations do not fit (K = 256).
// ------------------------------------------------------------------------------------------------
template <int NSPLIT>
__global__ void __launch_bounds__(kRowsThreads, 1)
point_mlp_rows_tc_kernel(const __grid_constant__ TcParams p, int num_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[2 * kMaxRing + 1];
    __shared__ uint32_t seq_off_s[kMaxSeq], seq_bytes_s[kMaxSeq];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, half = (tid >> 7) & 1;  // workers: tile row, which half of the columns
    constexpr int TR = kTileRows;
    constexpr uint32_t LBO = TR * 16;
    int kmax = 0, npmax = 0;
    for (int s = 0; s < p.na; s++) {
        kmax = max(kmax, p.a[s].Kp);
        npmax = max(npmax, p.a[s].Np);
    }
    const uint32_t x_img = (uint32_t)(kmax / 4) * LBO;
    uint8_t *x_hi = smem, *x_lo = smem + x_img;
    float *bias_s = reinterpret_cast<float *>(smem + (NSPLIT == 3 ? 2 : 1) * x_img);  // [na][npmax], zero padded
    Ring ring;
    ring.slots = reinterpret_cast<uint8_t *>(bias_s) + pad_to(p.na * npmax * 4, 128);
    ring.nslots = p.ring_slots;
    ring.full = bars;
    ring.empty = bars + kMaxRing;
    ring.reset();
    for (int i = 0; i < kMaxRing; i++) ring.off[i] = (uint32_t)i * kSlotBytes;
    uint64_t *bar_mma = bars + 2 * kMaxRing;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < ring.nslots; i++) {
            tc::mbar_init(&ring.full[i], 1);
            tc::mbar_init(&ring.empty[i], 1);
        }
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init_fence();
    }
    for (int i = tid; i < p.na * npmax; i += kRowsThreads) {
        const int s = i / npmax, j = i % npmax;
        bias_s[i] = j < p.a[s].Cout ? __ldg(p.a[s].bias + j) : 0.f;
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    SliceSeq prod;
    prod.n = p.na;
    prod.chunk_major = 0;
    for (int s = 0; s < p.na && s < 3; s++) prod.st[s] = &p.a[s];
    prod.reset();
    const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total_slices = (long long)my_tiles * prod.per_tile();
    const bool sticky = prod.per_tile() <= ring.nslots;
    if (sticky) ring.nslots = max(prod.per_tile(), 1);
    if (warp == 8 && lane == 0) {
        ring_start(ring, prod, seq_off_s, seq_bytes_s, p.packed, prod.per_tile(), total_slices, sticky);
    }
    uint32_t mma_phase = 0;
    const long long rows_total = (long long)p.c.B * p.c.Nprev;
    const int row_w = 4 + p.c.Cin;
    const uint32_t row_off = (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
    const int C = p.a[p.na - 1].Cout;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long row0 = (long long)tile * TR;
        // ---- X0 = table[row0 .. row0+128, 4:4+Cin] as hi/lo K-major images (coalesced 128-bit loads) ----
        if (warp < 8) {
            const int kp0 = p.a[0].Kp;
            for (int e = tid; e < TR * (kp0 / 4); e += 256) {
                const int r = e / (kp0 / 4), c = (e % (kp0 / 4)) * 4;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (row0 + r < rows_total) {
                    const float *src = p.c.table + (row0 + r) * row_w + 4 + c;
                    if ((row_w & 3) == 0 && c + 3 < p.c.Cin) {
                        float4 t4 = __ldg(reinterpret_cast<const float4 *>(src));
                        v[0] = t4.x; v[1] = t4.y; v[2] = t4.z; v[3] = t4.w;
                    } else {
                        for (int i = 0; i < 4; i++)
                            if (c + i < p.c.Cin) v[i] = __ldg(src + i);
                    }
                }
                float hi[4], lo[4];
                for (int i = 0; i < 4; i++) tc::split_op<NSPLIT>(v[i], hi[i], lo[i]);
                const uint32_t off = tc::kmajor_off(r, c, LBO);
                *reinterpret_cast<float4 *>(x_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                if (NSPLIT == 3) *reinterpret_cast<float4 *>(x_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        for (int s = 0; s < p.na; s++) {
            const TcStage &st = p.a[s];
            if (warp == 8) {
                if (lane == 0) {
                    run_transposed_stage<NSPLIT, true>(st.Kp, st.Np / 128, ring, sticky, tc::smem_u32(x_hi),
                                                       tc::smem_u32(x_lo), LBO, 128, tmem, 128);
                    tc::mma_commit(bar_mma);
                }
                __syncwarp();
            }
            wait_bar(bar_mma, mma_phase);
            mma_phase ^= 1;
            tc::fence_after_sync();
            if (warp < 8) {  // two threads per row: columns [0, mid) and [mid, width), mid a multiple of 16
                const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
                if (s + 1 < p.na) {
                    const int kpn = p.a[s + 1].Kp, mid = ((kpn + 31) / 32) * 16;
                    plain_epilogue<NSPLIT>(st, tl, row_off, x_hi, x_lo, LBO, kpn, bias_s + s * npmax,
                                           half ? mid : 0, half ? kpn : mid);
                } else if (row0 + row < rows_total) {  // last stage: F[row, :] = relu(D + b), 128-bit stores
                    float *dst = p.ftab + (row0 + row) * C;
                    const float *bs = bias_s + s * npmax;
                    const int mid = ((C + 31) / 32) * 16;
                    for (int c0 = half ? mid : 0; c0 < (half ? C : min(mid, C)); c0 += 16) {
                        uint32_t v[16];
                        tc::tmem_ld16(tl + c0, v);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int c = c0 + q * 4;
                            if (c + 3 < C) {
                                const float4 b4 = *reinterpret_cast<const float4 *>(bs + c);
                                *reinterpret_cast<float4 *>(dst + c) =
                                    make_float4(fmaxf(__uint_as_float(v[q * 4 + 0]) + b4.x, 0.f),
                                                fmaxf(__uint_as_float(v[q * 4 + 1]) + b4.y, 0.f),
                                                fmaxf(__uint_as_float(v[q * 4 + 2]) + b4.z, 0.f),
                                                fmaxf(__uint_as_float(v[q * 4 + 3]) + b4.w, 0.f));
                            } else {
                                for (int i = 0; i < 4; i++)
                                    if (c + i < C) dst[c + i] = fmaxf(__uint_as_float(v[q * 4 + i]) + bs[c + i], 0.f);
                            }
                        }
                    }
                }
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
            tc::fence_after_sync();
        }
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// Kernel B: per-edge attention MLP (+ the per-edge feature MLP of the first layer), product with
// the (gathered) features, max over the K slots, centre mask, output row.
//
// Per tile of 128 edges:
//   gather phase (thread = edge): neighbour xyz (128-bit load) -> geo / att_vec -> the tiny-K stages on
//     the CUDA cores in exact fp32 (attention stage 0, K <= 10; first-layer feature stage 0, K = 3)
//     -> hi/lo K-major operand images in shared memory;
//   first layer only: remaining hidden feature stages on the tensor core, D[edge, ch];
//   per 128-channel chunk: last feature stage (first layer) and attention stage 1 TRANSPOSED on the
//     tensor core, D^T[ch, edge]; epilogue thread = channel: relu(att) * feat (feat from TMEM for the
//     first layer, gathered from kernel A's table otherwise), running max over each centre's K
//     consecutive TMEM columns, pre-ReLU, centre mask, coalesced store.
// FIRST selects the first-layer variant (features computed per edge) at compile time so that neither
// variant carries the other's (predicated) instructions.
// ------------------------------------------------------------------------------------------------
template <int NSPLIT, bool FIRST>
__global__ void __launch_bounds__(kTcThreads, FIRST ? 2 : 3)
edge_tc_kernel(const __grid_constant__ TcParams p, int num_tiles, int cpt) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[2 * kMaxRing + 1];
    __shared__ uint32_t seq_off_s[kMaxSeq], seq_bytes_s[kMaxSeq];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t rowoff_s[kTileRows];  // element offset of every edge's source row in ftab
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ConvParams &c = p.c;
    const int K = c.K, C = c.Cout;
    constexpr uint32_t LBO = kTileRows * 16;  // images with 128 rows: panel = 2 KB
    constexpr int NIMG = NSPLIT == 3 ? 2 : 1;
    const bool has_att = p.has_att != 0;

    // ---- shared memory carve-up (mirrored by kernel_b_base on the host) ----
    int kx = 0;  // feature-path image width (first layer only)
    if (FIRST) {
        if (p.f0_cuda) kx = pad_to(p.f0_cout, 8);
        for (int s = 0; s < p.nfh; s++) kx = max(kx, max(p.fh[s].Kp, pad_to(p.fh[s].Cout, 8)));
        kx = max(kx, p.ff.Kp);
    }
    const int kh = has_att ? p.a1.Kp : 0;
    uint8_t *xf_hi = smem, *xf_lo = xf_hi + (size_t)(kx / 4) * LBO;
    uint8_t *xa_hi = smem + (size_t)NIMG * (kx / 4) * LBO, *xa_lo = xa_hi + (size_t)(kh / 4) * LBO;
    uint8_t *wres = xa_hi + (size_t)NIMG * (kh / 4) * LBO;  // resident plain-stage operand images
    size_t wres_bytes = 0;
    for (int s = 0; s < p.nfh; s++) wres_bytes += (size_t)2 * p.fh[s].Np * p.fh[s].Kp * 4;
    // CUDA-core stage weights, zero padded so that the inner loops need no bounds checks:
    //   wa0_s[kh][8]: folded attention stage-0 rows (att0_folded_weight; the region keeps its kh*12 floats);   wf0_s[kf0][4]: w0 w1 w2 bias
    float *wa0_s = reinterpret_cast<float *>(wres + wres_bytes);
    const int kf0 = (FIRST && p.f0_cuda) ? pad_to(p.f0_cout, 8) : 0;
    float *wf0_s = wa0_s + kh * 12;
    // biases: hidden feature stages (128 floats each, zero padded), then ff and a1 (Cp floats each)
    const int Cp = pad_to(C, 128);
    float *bias_s = wf0_s + kf0 * 4;
    float *bias_ff_s = bias_s + p.nfh * 128, *bias_a1_s = bias_ff_s + Cp;
    float *small_end = bias_a1_s + Cp;
    Ring ring;
    ring.slots = smem + pad_to((int)(reinterpret_cast<uint8_t *>(small_end) - smem), 128);
    ring.nslots = p.ring_slots;
    ring.full = bars;
    ring.empty = bars + kMaxRing;
    ring.reset();
    uint64_t *bar_mma = bars + 2 * kMaxRing;

    SliceSeq prod;
    prod.n = 0;
    prod.chunk_major = 1;
    if (FIRST) prod.st[prod.n++] = &p.ff;
    if (has_att) prod.st[prod.n++] = &p.a1;
    prod.reset();
    const int per_tile = prod.per_tile();
    const bool sticky = p.ring_sticky != 0;
    if (sticky) {  // exact-fit slots in sequence order
        SliceSeq tmp = prod;
        uint32_t o = 0;
        for (int i = 0; i < per_tile && i < kMaxRing; i++) {
            long long off;
            uint32_t bytes;
            tmp.next(off, bytes);
            ring.off[i] = o;
            o += bytes;
        }
    } else {
        for (int i = 0; i < kMaxRing; i++) ring.off[i] = (uint32_t)i * kSlotBytes;
    }

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < ring.nslots; i++) {
            tc::mbar_init(&ring.full[i], 1);
            tc::mbar_init(&ring.empty[i], 1);
        }
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init_fence();
    }
    // resident operands: tensor-core images of the plain stages, fp32 weights of the CUDA-core stages
    {
        size_t off = 0;
        for (int s = 0; s < p.nfh; s++) {
            const TcStage &st = p.fh[s];
            const int n4 = 2 * st.Np * st.Kp / 4;
            const float4 *src = reinterpret_cast<const float4 *>(p.packed + st.w_off);
            float4 *dst = reinterpret_cast<float4 *>(wres + off);
            for (int i = tid; i < n4; i += kTcThreads) dst[i] = __ldg(src + i);
            off += (size_t)n4 * 16;
        }
        for (int i = tid; i < kh * 8; i += kTcThreads)  // folded rows, see att0_folded_weight
            wa0_s[i] = att0_folded_weight(p.a0_w, p.a0_b, p.a0_cin, p.a0_cout, i >> 3, i & 7);
        for (int i = tid; i < kf0 * 4; i += kTcThreads) {
            const int j = i / 4, q = i % 4;
            wf0_s[i] = j < p.f0_cout ? (q < 3 ? __ldg(p.f0_w + (size_t)j * 3 + q) : __ldg(p.f0_b + j)) : 0.f;
        }
        for (int i = tid; i < p.nfh * 128; i += kTcThreads) {
            const int st = i >> 7, j = i & 127;
            bias_s[i] = j < p.fh[st].Cout ? __ldg(p.fh[st].bias + j) : 0.f;
        }
        for (int i = tid; i < Cp; i += kTcThreads) {
            bias_ff_s[i] = (FIRST && i < C) ? __ldg(p.ff.bias + i) : 0.f;
            bias_a1_s[i] = (has_att && i < C) ? __ldg(p.a1.bias + i) : 0.f;
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total_slices = (long long)my_tiles * per_tile;
    if (warp == 4 && lane == 0) {
        ring_start(ring, prod, seq_off_s, seq_bytes_s, p.packed, per_tile, total_slices, sticky);
    }
    uint32_t mma_phase = 0;
    const long long rows_total = (long long)c.B * c.Nprev;
    const long long centers_total = (long long)c.B * c.O;
    const int row_w = 4 + c.Cin;
    const int out_w = 4 + C;
    const int attfdim = c.attfdim, Nprev = c.Nprev, O = c.O;
    const uint32_t tm_f = tmem, tm_g = tmem + (FIRST ? 128 : 0);
    const uint32_t row_off = (uint32_t)(tid >> 3) * 128u + (uint32_t)(tid & 7) * 16u;  // this edge's row in an image
    const int nchunk = pad_to(C, 128) / 128;

    // optional per-phase cycle accounting of CTA 0 / thread 0 (debug)
    long long tph = 0;
    unsigned long long acc_ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool timing = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
#define GG_PHASE(i)                                   \
    if (timing) {                                     \
        const long long now_ = clock64();             \
        acc_ph[i] += (unsigned long long)(now_ - tph); \
        tph = now_;                                   \
    }
    // per-edge prefetch state (thread = edge row of the tile): the neighbour index of the NEXT tile is
    // requested during this tile's gather phase, its table row and centre under this tile's MMAs.
    const int my_cl = tid / K, my_slot = tid - my_cl * K;
    bool pf_valid = false;
    int pf_idx = 0;
    long long pf_center = 0;
    uint32_t pf_roff = 0;
    float4 pf_head = make_float4(0.f, 0.f, 0.f, 0.f), pf_cent = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch_row = [&]() {  // table row + centre of the edge whose index is in pf_idx
        pf_roff = 0;
        if (pf_valid) {
            const int b = (int)(pf_center / O);
            const long long row = take_row(pf_idx, b, Nprev, rows_total);
            pf_roff = (uint32_t)row * (uint32_t)C;
            const float *src = c.table + row * row_w;
            if ((row_w & 3) == 0) {
                pf_head = __ldg(reinterpret_cast<const float4 *>(src));
            } else {
                pf_head = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 0.f);
            }
            pf_cent = __ldg(c.cent + pf_center);
        }
    };
    if (warp < 4 && blockIdx.x < num_tiles) {
        const long long center = (long long)blockIdx.x * cpt + my_cl;
        pf_valid = my_cl < cpt && center < centers_total;
        pf_center = center;
        if (pf_valid) pf_idx = __ldg(c.nebidx + center * K + my_slot);
        fetch_row();
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long c_base = (long long)tile * cpt;
        if (timing) tph = clock64();
        // ---- gather + geometry + tiny-K stages on the CUDA cores: thread = edge row ----
        if (warp < 4) {
            float att[12];
#pragma unroll
            for (int i = 0; i < 12; i++) att[i] = 0.f;
            att[10] = 1.f;  // multiplies the bias column of wa0_s
            float dx = 0.f, dy = 0.f, dz = 0.f;
            const uint32_t roff = pf_roff;
            if (pf_valid) att_vector(attfdim, pf_cent, pf_head.x, pf_head.y, pf_head.z, att, dx, dy, dz);
            const Att7 a7 = fold_att(attfdim, att);
            {   // request the next tile's neighbour index now; its row / centre are requested later
                const long long ncenter = (long long)(tile + gridDim.x) * cpt + my_cl;
                pf_valid = (tile + (int)gridDim.x) < num_tiles && my_cl < cpt && ncenter < centers_total;
                pf_center = ncenter;
                if (pf_valid) pf_idx = __ldg(c.nebidx + ncenter * K + my_slot);
            }
            if (!FIRST) rowoff_s[tid] = roff;
            // attention stage 0: h = relu(W a + b), K <= 10, exact fp32 (bias folded in as w[10] * 1)
            for (int g = 0; g < kh / 4; g++) {
                const float4 *w4 = reinterpret_cast<const float4 *>(wa0_s + g * 32);
                float hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float acc = att0_channel(w4[i * 2], w4[i * 2 + 1], a7);
                    tc::split_op<NSPLIT>(fmaxf(acc, 0.f), hi[i], lo[i]);
                }
                const uint32_t off = row_off + (uint32_t)g * LBO;
                *reinterpret_cast<float4 *>(xa_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                if (NSPLIT == 3)
                    *reinterpret_cast<float4 *>(xa_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            if (FIRST) {  // features start from the geo vector (gcn_module_g_att.py:242-243)
                if (kf0 > 0) {  // feature stage 0 (K = 3) on the CUDA cores
                    for (int g = 0; g < kf0 / 4; g++) {
                        const float4 *w4 = reinterpret_cast<const float4 *>(wf0_s + g * 16);
                        float hi[4], lo[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const float4 w = w4[i];
                            const float v = fmaf(w.z, dz, fmaf(w.y, dy, fmaf(w.x, dx, w.w)));
                            tc::split_op<NSPLIT>(fmaxf(v, 0.f), hi[i], lo[i]);
                        }
                        const uint32_t off = row_off + (uint32_t)g * LBO;
                        *reinterpret_cast<float4 *>(xf_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        if (NSPLIT == 3)
                            *reinterpret_cast<float4 *>(xf_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                } else {  // single-stage feature MLP: the tensor core reads the geo vector itself (K = 8)
                    float hi[4], lo[4];
                    tc::split_op<NSPLIT>(dx, hi[0], lo[0]);
                    tc::split_op<NSPLIT>(dy, hi[1], lo[1]);
                    tc::split_op<NSPLIT>(dz, hi[2], lo[2]);
                    hi[3] = lo[3] = 0.f;
                    *reinterpret_cast<float4 *>(xf_hi + row_off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4 *>(xf_hi + row_off + LBO) = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (NSPLIT == 3) {
                        *reinterpret_cast<float4 *>(xf_lo + row_off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                        *reinterpret_cast<float4 *>(xf_lo + row_off + LBO) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
        }
        GG_PHASE(0)  // gather + CUDA-core stages
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        GG_PHASE(1)  // sync after gather

        // ---- first layer: remaining hidden feature stages on the tensor core, D[edge, ch] ----
        if (FIRST) {
            size_t woff = 0;
            for (int s = 0; s < p.nfh; s++) {
                const TcStage &st = p.fh[s];
                if (warp == 4) {
                    if (lane == 0) {
                        run_plain_stage<NSPLIT>(st, tc::smem_u32(xf_hi), tc::smem_u32(xf_lo), LBO,
                                                tc::smem_u32(wres + woff), tmem);
                        tc::mma_commit(bar_mma);
                    }
                    __syncwarp();
                }
                woff += (size_t)2 * st.Np * st.Kp * 4;
                wait_bar(bar_mma, mma_phase);
                mma_phase ^= 1;
                tc::fence_after_sync();
                GG_PHASE(2)  // hidden-stage MMA issue -> retired
                if (warp < 4) {
                    const int kp_next = s + 1 < p.nfh ? p.fh[s + 1].Kp : p.ff.Kp;
                    plain_epilogue<NSPLIT>(st, tmem + ((uint32_t)(warp * 32) << 16), row_off, xf_hi, xf_lo, LBO, kp_next,
                                           bias_s + s * 128);
                }
                GG_PHASE(3)  // hidden-stage epilogue
                tc::fence_async_smem();
                tc::fence_before_sync();
                __syncthreads();
                tc::fence_after_sync();
                GG_PHASE(4)  // sync after hidden stage
            }
        }

        // ---- transposed stages, one 128-channel chunk at a time ----
        for (int j = 0; j < nchunk; j++) {
            if (warp == 4) {
                if (lane == 0) {
                    if (FIRST) {
                        run_transposed_stage<NSPLIT>(p.ff.Kp, 1, ring, sticky, tc::smem_u32(xf_hi),
                                                     tc::smem_u32(xf_lo), LBO, 128, tm_f, 0);
                    }
                    if (has_att) {
                        run_transposed_stage<NSPLIT>(p.a1.Kp, 1, ring, sticky, tc::smem_u32(xa_hi),
                                                     tc::smem_u32(xa_lo), LBO, 128, tm_g, 0);
                    }
                    tc::mma_commit(bar_mma);
                }
                __syncwarp();
            }
            // ---- final epilogue: thread = channel (TMEM lane), 16 edge columns per batch, the TMEM loads (and
            //      feature gathers) of batch k+1 in flight while batch k is reduced.  Warps whose 32 channels
            //      all lie beyond C skip it entirely. ----
            const int ch = j * 128 + tid;
            const bool chv = ch < C;
            const bool warp_on = warp < 4 && (j * 128 + warp * 32) < C;
            // gathered feature of edge e for this channel: fbase[rowoff_s[e]] (invalid rows / channels are
            // clamped to a valid address; their values never reach the output)
            const float *fbase = FIRST ? nullptr : p.ftab + (chv ? ch : 0);
            uint32_t gA[16], fA[16], gB[16], fB[16];  // f*: TMEM bits (first layer) or gathered floats
            auto gather16 = [&](uint32_t (&f)[16], int e0) {
#pragma unroll
                for (int i = 0; i < 16; i++) f[i] = __float_as_uint(__ldg(fbase + rowoff_s[e0 + i]));
            };
            if (!FIRST && warp_on) gather16(fA, 0);  // first batch of feature gathers issued under the MMAs
            if (j == 0 && warp < 4) fetch_row();      // next tile's table row + centre, also under the MMAs
            wait_bar(bar_mma, mma_phase);
            mma_phase ^= 1;
            tc::fence_after_sync();
            GG_PHASE(5)  // transposed-stage MMAs -> retired
            if (warp_on) {
                const float bf = FIRST ? bias_ff_s[chv ? ch : 0] : 0.f;
                const float ba = has_att ? bias_a1_s[chv ? ch : 0] : 0.f;
                const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
                const float pre_floor = c.pre_relu ? 0.f : -3.402823466e+38f;
                float *out_ch = c.out + 4 + ch;
                float m = -3.402823466e+38f;
                int pos = 0, cl = 0;
                auto issue = [&](uint32_t (&g)[16], uint32_t (&f)[16], int e0) {
                    if (has_att) tc::tmem_ld16(tm_g + lane_off + e0, g);
                    if (FIRST) tc::tmem_ld16(tm_f + lane_off + e0, f);
                };
                auto reduce16 = [&](const uint32_t (&g)[16], const uint32_t (&f)[16]) {
                    float pr[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        float x = FIRST ? fmaxf(__uint_as_float(f[i]) + bf, 0.f) : __uint_as_float(f[i]);
                        if (has_att) x *= fmaxf(__uint_as_float(g[i]) + ba, 0.f);  // :167 att * feats
                        pr[i] = x;
                    }
                    if ((K & 15) == 0) {  // the 16 columns belong to one centre
                        float mm = pr[0];
#pragma unroll
                        for (int i = 1; i < 16; i++) mm = fmaxf(mm, pr[i]);
                        m = fmaxf(m, mm);
                        pos += 16;
                        if (pos == K) {
                            const long long center = c_base + cl;
                            if (chv && cl < cpt && center < centers_total)
                                out_ch[center * out_w] = fmaxf(m, pre_floor) * __ldg(c.centmsk + center);
                            pos = 0;
                            cl++;
                            m = -3.402823466e+38f;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            m = fmaxf(m, pr[i]);
                            if (++pos == K) {
                                const long long center = c_base + cl;
                                if (chv && cl < cpt && center < centers_total)
                                    out_ch[center * out_w] = fmaxf(m, pre_floor) * __ldg(c.centmsk + center);
                                pos = 0;
                                cl++;
                                m = -3.402823466e+38f;
                            }
                        }
                    }
                };
                issue(gA, fA, 0);
                tc::tmem_ld_wait();
#pragma unroll 1
                for (int e0 = 0; e0 < kTileRows; e0 += 32) {
                    issue(gB, fB, e0 + 16);
                    if (!FIRST) gather16(fB, e0 + 16);
                    reduce16(gA, fA);
                    tc::tmem_ld_wait();
                    if (e0 + 32 < kTileRows) {
                        issue(gA, fA, e0 + 32);
                        if (!FIRST) gather16(fA, e0 + 32);
                    }
                    reduce16(gB, fB);
                    tc::tmem_ld_wait();
                }
            }
            GG_PHASE(6)  // final epilogue
            tc::fence_before_sync();
            __syncthreads();
            tc::fence_after_sync();
            GG_PHASE(7)  // sync after final epilogue
        }
        // centre columns of the output rows
        if (warp < 4) {
            for (int i = tid; i < cpt * 4; i += 128) {
                const long long center = c_base + i / 4;
                if (center < centers_total)
                    c.out[center * out_w + (i & 3)] =
                        __ldg(reinterpret_cast<const float *>(c.cent + center) + (i & 3));
            }
        }
    }
    if (timing)
        for (int i = 0; i < 8; i++) p.dbg[i] = acc_ph[i];
#undef GG_PHASE
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// Kernel B', compact first-layer variant (Cout <= 64, K | 64 or K = 128, 3-pass mode): same math as
// edge_tc_kernel<3, true>, re-laid-out so that THREE CTAs fit on an SM (the first layer is latency bound:
// more resident tiles is what speeds it up; profiles/README.md).
//   * M = 64 MMAs: the 64 feature channels F and the 64 attention channels G share ONE 128-column TMEM
//     tile.  A cta_group::1 M=64 accumulator occupies lanes 32q..32q+15 of every quadrant q (row 16q+i ->
//     lane 32q+i); issued with lane offset 16 it occupies lanes 32q+16..32q+31 (tools/m64_probe.cu).  F goes
//     to the low half-warps, G to the high ones: lane l and lane l^16 hold the same channel, the product
//     needs one SHFL, and all four epilogue warps are busy (the 128-row variant idles two).
//   * TMEM: 128 columns (F|G) + a second allocation for the hidden-stage accumulators (32): 160 <= 512/3.
//   * shared memory: the attention operand image aliases the lo half of the feature image (attention stage 1
//     is issued first and has retired before the feature path writes its lo parts -- which wait in the idle
//     hidden accumulator's TMEM columns of their thread's lane meanwhile, tcgen05.st / ld), and the last-stage
//     weight images keep only their 64 live rows: ~66 KB per CTA instead of ~108 KB.
//   Tried and dropped (r01, measured): starting the F|G accumulators at the bias through tcgen05.st (saves the
//     bias adds, but the stores + wait::st sit on the tile's critical path: +6 %); picking the epilogue halves
//     by TMEM address (tcgen05.ld addresses must be warp-uniform).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc_raw(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

struct First64Layout {  // byte offsets inside dynamic shared memory (host and device agree through this)
    int kx, kh, kf0;
    uint32_t xf_lo, wres, wff_hi, wff_lo, wa1_hi, wa1_lo, wa0, wf0, bias, total;
};
__host__ __device__ inline First64Layout first64_layout(const TcParams &p) {
    First64Layout L;
    int kx = pad_to(p.f0_cout, 8);
    for (int s = 0; s < p.nfh; s++) kx = max(kx, max(p.fh[s].Kp, pad_to(p.fh[s].Cout, 8)));
    kx = max(kx, p.ff.Kp);
    L.kx = kx;
    L.kh = p.a1.Kp;
    L.kf0 = pad_to(p.f0_cout, 8);
    L.xf_lo = (uint32_t)(kx / 4) * 2048u;
    L.wres = 2u * L.xf_lo;
    uint32_t o = L.wres;
    for (int s = 0; s < p.nfh; s++) o += 2u * p.fh[s].Np * p.fh[s].Kp * 4u;
    L.wff_hi = o;
    L.wff_lo = L.wff_hi + (uint32_t)(p.ff.Kp / 4) * 1024u;
    L.wa1_hi = L.wff_lo + (uint32_t)(p.ff.Kp / 4) * 1024u;
    L.wa1_lo = L.wa1_hi + (uint32_t)(L.kh / 4) * 1024u;
    L.wa0 = L.wa1_lo + (uint32_t)(L.kh / 4) * 1024u;
    L.wf0 = L.wa0 + (uint32_t)L.kh * 12u * 4u;
    L.bias = L.wf0 + (uint32_t)L.kf0 * 4u * 4u;
    L.total = L.bias + (uint32_t)p.nfh * 128u * 4u;
    return L;
}

template <int NSPLIT>
__global__ void __launch_bounds__(kTcThreads, 3)
edge_first64_kernel(const __grid_constant__ TcParams p, int num_tiles, int cpt, int hid_cols, int stash_lo) {
    static_assert(NSPLIT == 3, "compact first-layer kernel: 3-pass mode only");
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_mma_s;
    __shared__ uint32_t tmem_base_s[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ConvParams &c = p.c;
    const int K = c.K, C = c.Cout;
    constexpr uint32_t LBO = kTileRows * 16;  // 128-row images: panel = 2 KB
    constexpr uint32_t WLBO = 64 * 16;        // 64-row weight images: panel = 1 KB
    const First64Layout L = first64_layout(p);
    const int kh = L.kh, kf0 = L.kf0;
    uint8_t *xf_hi = smem, *xf_lo = smem + L.xf_lo;
    uint8_t *xa_hi = xf_lo, *xa_lo = xa_hi + (size_t)(kh / 4) * LBO;  // alias (2 * kh <= kx, host-checked)
    uint8_t *wres = smem + L.wres;
    float *wa0_s = reinterpret_cast<float *>(smem + L.wa0);
    float *wf0_s = reinterpret_cast<float *>(smem + L.wf0);
    float *bias_s = reinterpret_cast<float *>(smem + L.bias);
    uint64_t *bar_mma = &bar_mma_s;

    if (warp == 0) {
        tmem_alloc_raw(&tmem_base_s[0], 128);
        tmem_alloc_raw(&tmem_base_s[1], (uint32_t)hid_cols);
        tmem_relinquish();
    }
    if (tid == 0) {
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init_fence();
    }
    {   // resident operands
        size_t off = 0;
        for (int s = 0; s < p.nfh; s++) {
            const TcStage &st = p.fh[s];
            const int n4 = 2 * st.Np * st.Kp / 4;
            const float4 *src = reinterpret_cast<const float4 *>(p.packed + st.w_off);
            float4 *dst = reinterpret_cast<float4 *>(wres + off);
            for (int i = tid; i < n4; i += kTcThreads) dst[i] = __ldg(src + i);
            off += (size_t)n4 * 16;
        }
        // rows 0..63 of every [128 x 4] panel of the packed transposed-stage images (chunk 0)
        auto load_compact = [&](const TcStage &st, uint32_t dst_hi, uint32_t dst_lo) {
            const int per_img = (st.Kp / 4) * 256;
            for (int i = tid; i < 2 * per_img; i += kTcThreads) {
                const int img = i / per_img, r = i % per_img, P = r >> 8, w = r & 255;
                const int t = P / (kSliceK / 4), pp = P % (kSliceK / 4);
                const int kw = min(kSliceK, st.Kp - t * kSliceK);
                const float *src = p.packed + st.w_off + (size_t)t * 2 * 128 * kSliceK + (img ? 128 * kw : 0) +
                                   pp * 512 + w;
                reinterpret_cast<float *>(smem + (img ? dst_lo : dst_hi))[r] = __ldg(src);
            }
        };
        load_compact(p.ff, L.wff_hi, L.wff_lo);
        load_compact(p.a1, L.wa1_hi, L.wa1_lo);
        for (int i = tid; i < kh * 8; i += kTcThreads)  // folded rows, see att0_folded_weight
            wa0_s[i] = att0_folded_weight(p.a0_w, p.a0_b, p.a0_cin, p.a0_cout, i >> 3, i & 7);
        for (int i = tid; i < kf0 * 4; i += kTcThreads) {
            const int j = i / 4, q = i % 4;
            wf0_s[i] = j < p.f0_cout ? (q < 3 ? __ldg(p.f0_w + (size_t)j * 3 + q) : __ldg(p.f0_b + j)) : 0.f;
        }
        for (int i = tid; i < p.nfh * 128; i += kTcThreads) {
            const int st = i >> 7, j = i & 127;
            bias_s[i] = j < p.fh[st].Cout ? __ldg(p.fh[st].bias + j) : 0.f;
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_fg = tmem_base_s[0], tmem_h = tmem_base_s[1];

    uint32_t mma_phase = 0;
    const long long rows_total = (long long)c.B * c.Nprev;
    const long long centers_total = (long long)c.B * c.O;
    const int row_w = 4 + c.Cin;
    const int out_w = 4 + C;
    const int attfdim = c.attfdim, Nprev = c.Nprev, O = c.O;
    const uint32_t row_off = (uint32_t)(tid >> 3) * 128u + (uint32_t)(tid & 7) * 16u;
    // final epilogue role of this thread: TMEM lane = 32*warp + lane; low half-warp: feature channel, high: attention
    const bool is_g = lane >= 16;
    const int ch = 16 * (warp & 3) + (lane & 15);
    const bool chv = warp < 4 && ch < C;
    const float bias_fg = chv ? __ldg((is_g ? p.a1.bias : p.ff.bias) + ch) : 0.f;
    const float pre_floor = c.pre_relu ? 0.f : -3.402823466e+38f;
    const uint32_t lane_taddr = ((uint32_t)((warp & 3) * 32)) << 16;  // this thread's TMEM lane quadrant

    const int my_cl = tid / K, my_slot = tid - my_cl * K;
    bool pf_valid = false;
    int pf_idx = 0;
    long long pf_center = 0;
    float4 pf_head = make_float4(0.f, 0.f, 0.f, 0.f), pf_cent = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch_row = [&]() {
        if (pf_valid) {
            const int b = (int)(pf_center / O);
            const long long row = take_row(pf_idx, b, Nprev, rows_total);
            const float *src = c.table + row * row_w;
            if ((row_w & 3) == 0) pf_head = __ldg(reinterpret_cast<const float4 *>(src));
            else pf_head = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 0.f);
            pf_cent = __ldg(c.cent + pf_center);
        }
    };
    if (warp < 4 && blockIdx.x < num_tiles) {
        const long long center = (long long)blockIdx.x * cpt + my_cl;
        pf_valid = my_cl < cpt && center < centers_total;
        pf_center = center;
        if (pf_valid) pf_idx = __ldg(c.nebidx + center * K + my_slot);
        fetch_row();
    }
    const uint32_t idesc64 = tc::make_idesc_tf32(64, kTileRows);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long c_base = (long long)tile * cpt;
        float dx = 0.f, dy = 0.f, dz = 0.f;
        // ---- gather + attention stage 0 (thread = edge row) ----
        if (warp < 4) {
            float att[12];
#pragma unroll
            for (int i = 0; i < 12; i++) att[i] = 0.f;
            att[10] = 1.f;
            if (pf_valid) att_vector(attfdim, pf_cent, pf_head.x, pf_head.y, pf_head.z, att, dx, dy, dz);
            const Att7 a7 = fold_att(attfdim, att);
            {
                const long long ncenter = (long long)(tile + gridDim.x) * cpt + my_cl;
                pf_valid = (tile + (int)gridDim.x) < num_tiles && my_cl < cpt && ncenter < centers_total;
                pf_center = ncenter;
                if (pf_valid) pf_idx = __ldg(c.nebidx + ncenter * K + my_slot);
            }
            for (int g = 0; g < kh / 4; g++) {
                const float4 *w4 = reinterpret_cast<const float4 *>(wa0_s + g * 32);
                float hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float acc = att0_channel(w4[i * 2], w4[i * 2 + 1], a7);
                    tc::split_op<NSPLIT>(fmaxf(acc, 0.f), hi[i], lo[i]);
                }
                const uint32_t off = row_off + (uint32_t)g * LBO;
                *reinterpret_cast<float4 *>(xa_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(xa_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        // ---- attention stage 1 -> G (lanes 32q+16..31), issued now, needed only by the final epilogue ----
        if (warp == 4) {
            if (lane == 0) {
                const uint32_t d = tmem_fg + (16u << 16);
                uint32_t acc = 0;
                for (int ks = 0; ks < kh / 8; ks++) {
                    const uint64_t ah = tc::make_sdesc(tc::smem_u32(smem + L.wa1_hi) + ks * 2 * WLBO, WLBO);
                    const uint64_t al = tc::make_sdesc(tc::smem_u32(smem + L.wa1_lo) + ks * 2 * WLBO, WLBO);
                    const uint64_t bh = tc::make_sdesc(tc::smem_u32(xa_hi) + ks * 2 * LBO, LBO);
                    const uint64_t bl = tc::make_sdesc(tc::smem_u32(xa_lo) + ks * 2 * LBO, LBO);
                    tc::mma_tf32(d, al, bh, idesc64, acc);
                    tc::mma_tf32(d, ah, bl, idesc64, 1);
                    tc::mma_tf32(d, ah, bh, idesc64, 1);
                    acc = 1;
                }
                tc::mma_commit(bar_mma);
            }
            __syncwarp();
        }
        // ---- feature stage 0 (K = 3) on the CUDA cores: hi parts while the attention MMAs run, lo parts
        //      (whose image the attention operand aliases) once they have retired ----
        if (warp < 4) {
            for (int g = 0; g < kf0 / 4; g++) {
                const float4 *w4 = reinterpret_cast<const float4 *>(wf0_s + g * 16);
                float v[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 w = w4[i];
                    v[i] = fmaxf(fmaf(w.z, dz, fmaf(w.y, dy, fmaf(w.x, dx, w.w))), 0.f);
                }
                *reinterpret_cast<float4 *>(xf_hi + row_off + (uint32_t)g * LBO) = make_float4(v[0], v[1], v[2], v[3]);
                if (stash_lo) {  // the lo parts wait in this thread's lane of the (idle) hidden accumulator
                    float h0, h1, h2, h3, l0, l1, l2, l3;
                    tc::split_op<NSPLIT>(v[0], h0, l0); tc::split_op<NSPLIT>(v[1], h1, l1);
                    tc::split_op<NSPLIT>(v[2], h2, l2); tc::split_op<NSPLIT>(v[3], h3, l3);
                    tc::tmem_st4(tmem_h + lane_taddr + (uint32_t)(g * 4), l0, l1, l2, l3);
                }
            }
            if (stash_lo) tc::tmem_st_wait();
        }
        wait_bar(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc::fence_after_sync();
        if (warp < 4 && stash_lo) {
            for (int g = 0; g < kf0 / 4; g += 4) {  // kf0 is a multiple of 8: groups of up to 4 x 4 columns
                uint32_t l[4][4];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (g + q < kf0 / 4) tc::tmem_ld4(tmem_h + lane_taddr + (uint32_t)((g + q) * 4), l[q]);
                tc::tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (g + q < kf0 / 4)
                        *reinterpret_cast<uint4 *>(xf_lo + row_off + (uint32_t)(g + q) * LBO) =
                            make_uint4(l[q][0], l[q][1], l[q][2], l[q][3]);
            }
        } else if (warp < 4) {  // no room in the hidden accumulator: recompute
            for (int g = 0; g < kf0 / 4; g++) {
                const float4 *w4 = reinterpret_cast<const float4 *>(wf0_s + g * 16);
                float lo[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 w = w4[i];
                    float hi;
                    tc::split_op<NSPLIT>(fmaxf(fmaf(w.z, dz, fmaf(w.y, dy, fmaf(w.x, dx, w.w))), 0.f), hi, lo[i]);
                }
                *reinterpret_cast<float4 *>(xf_lo + row_off + (uint32_t)g * LBO) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        // ---- hidden feature stages on the tensor core, D[edge, ch] in the second TMEM allocation ----
        {
            size_t woff = 0;
            for (int s = 0; s < p.nfh; s++) {
                const TcStage &st = p.fh[s];
                if (warp == 4) {
                    if (lane == 0) {
                        run_plain_stage<NSPLIT>(st, tc::smem_u32(xf_hi), tc::smem_u32(xf_lo), LBO,
                                                tc::smem_u32(wres + woff), tmem_h);
                        tc::mma_commit(bar_mma);
                    }
                    __syncwarp();
                }
                woff += (size_t)2 * st.Np * st.Kp * 4;
                wait_bar(bar_mma, mma_phase);
                mma_phase ^= 1;
                tc::fence_after_sync();
                if (warp < 4) {
                    const int kp_next = s + 1 < p.nfh ? p.fh[s + 1].Kp : p.ff.Kp;
                    plain_epilogue<NSPLIT>(st, tmem_h + ((uint32_t)(warp * 32) << 16), row_off, xf_hi, xf_lo, LBO,
                                           kp_next, bias_s + s * 128);
                }
                tc::fence_async_smem();
                tc::fence_before_sync();
                __syncthreads();
                tc::fence_after_sync();
            }
        }
        // ---- last feature stage -> F (lanes 32q..32q+15) ----
        if (warp == 4) {
            if (lane == 0) {
                uint32_t acc = 0;
                for (int ks = 0; ks < p.ff.Kp / 8; ks++) {
                    const uint64_t ah = tc::make_sdesc(tc::smem_u32(smem + L.wff_hi) + ks * 2 * WLBO, WLBO);
                    const uint64_t al = tc::make_sdesc(tc::smem_u32(smem + L.wff_lo) + ks * 2 * WLBO, WLBO);
                    const uint64_t bh = tc::make_sdesc(tc::smem_u32(xf_hi) + ks * 2 * LBO, LBO);
                    const uint64_t bl = tc::make_sdesc(tc::smem_u32(xf_lo) + ks * 2 * LBO, LBO);
                    tc::mma_tf32(tmem_fg, al, bh, idesc64, acc);
                    tc::mma_tf32(tmem_fg, ah, bl, idesc64, 1);
                    tc::mma_tf32(tmem_fg, ah, bh, idesc64, 1);
                    acc = 1;
                }
                tc::mma_commit(bar_mma);
            }
            __syncwarp();
        }
        if (warp < 4) fetch_row();  // next tile's table row + centre, under the MMAs
        wait_bar(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc::fence_after_sync();
        // ---- final epilogue: lane pair (l, l^16) = (feature, attention) of one channel.  The low half-warp
        //      reduces columns [0, 64), the high one columns [64, 128): each sends the other the half it does
        //      not reduce, so one shuffle serves two columns.  K | 64: no centre straddles column 64. ----
        if (warp < 4) {
            const uint32_t taddr = tmem_fg + ((uint32_t)(warp * 32) << 16);
            float *out_ch = c.out + 4 + ch;
            float m = -3.402823466e+38f;
            int pos = 0, cl = is_g ? (cpt >> 1) : 0;
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t a[16], b[16];
                tc::tmem_ld16(taddr + c0, a);
                tc::tmem_ld16(taddr + 64 + c0, b);
                tc::tmem_ld_wait();
                float pr[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float xa = fmaxf(__uint_as_float(a[i]) + bias_fg, 0.f);
                    const float xb = fmaxf(__uint_as_float(b[i]) + bias_fg, 0.f);
                    const float got = __shfl_xor_sync(0xffffffffu, is_g ? xa : xb, 16);
                    pr[i] = (is_g ? xb : xa) * got;  // :167 att * feats
                }
                if ((K & 15) == 0) {
                    float mm = pr[0];
#pragma unroll
                    for (int i = 1; i < 16; i++) mm = fmaxf(mm, pr[i]);
                    m = fmaxf(m, mm);
                    pos += 16;
                    if (pos == K) {
                        const long long center = c_base + cl;
                        if (chv && center < centers_total)
                            out_ch[center * out_w] = fmaxf(m, pre_floor) * __ldg(c.centmsk + center);
                        pos = 0;
                        cl++;
                        m = -3.402823466e+38f;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        m = fmaxf(m, pr[i]);
                        if (++pos == K) {
                            const long long center = c_base + cl;
                            if (chv && center < centers_total)
                                out_ch[center * out_w] = fmaxf(m, pre_floor) * __ldg(c.centmsk + center);
                            pos = 0;
                            cl++;
                            m = -3.402823466e+38f;
                        }
                    }
                }
            }
            if (K == 2 * 64) {  // one centre per tile: its two 64-column halves meet across the lane pair
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
                if (!is_g && chv && c_base < centers_total)
                    out_ch[c_base * out_w] = fmaxf(m, pre_floor) * __ldg(c.centmsk + c_base);
            }
            for (int i = tid; i < cpt * 4; i += 128) {  // centre columns of the output rows
                const long long center = c_base + i / 4;
                if (center < centers_total)
                    c.out[center * out_w + (i & 3)] = __ldg(reinterpret_cast<const float *>(c.cent + center) + (i & 3));
            }
        }
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    __syncthreads();
    if (warp == 0) {
        tc::tmem_dealloc(tmem_fg, 128);
        tc::tmem_dealloc(tmem_h, (uint32_t)hid_cols);
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel B'', wide-layer variant (layers >= 1 whose operand image + weights leave room for ONE CTA per SM):
// the same tile pipeline as edge_tc_kernel<NSPLIT, false> with EIGHT worker warps.  With one resident CTA
// the four-warp version leaves the SM latency bound (issue slots ~20 % busy); here two threads share an
// edge row in the gather phase (each computes half of the attention stage-0 channels) and two warps share
// a TMEM lane quadrant in the epilogue (columns [0,64) / [64,128); K | 64 so no centre straddles the
// split).  Every 128-channel chunk has its own TMEM region and "retired" barrier: the MMAs of all chunks of
// a tile are issued back to back and the epilogue of chunk j overlaps the tensor-core work on the later
// chunks.  Shared-memory layout and ring protocol are those of edge_tc_kernel (kernel_b_base).
// ------------------------------------------------------------------------------------------------
constexpr int kWideThreads = 288;  // warps 0-7: workers; warp 8: TMA + MMA issue

template <int NSPLIT>
__global__ void __launch_bounds__(kWideThreads, 1)
edge_wide_kernel(const __grid_constant__ TcParams p, int num_tiles, int cpt) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int kWideSeq = 32, kWideChunks = 4;       // slices per tile (<= 4 chunks x <= 8 slices), chunks
    __shared__ uint64_t bars[2 * kMaxRing + kWideChunks];  // ring full / empty, one "MMAs retired" barrier per chunk
    __shared__ uint32_t seq_off_s[kWideSeq], seq_bytes_s[kWideSeq];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t rowoff_s[kTileRows];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, half = (tid >> 7) & 1;  // workers: edge row of the tile, which half of the work
    const ConvParams &c = p.c;
    const int K = c.K, C = c.Cout;
    constexpr uint32_t LBO = kTileRows * 16;
    constexpr int NIMG = NSPLIT == 3 ? 2 : 1;
    const int kh = p.a1.Kp;
    uint8_t *xa_hi = smem, *xa_lo = xa_hi + (size_t)(kh / 4) * LBO;
    float *wa0_s = reinterpret_cast<float *>(smem + (size_t)NIMG * (kh / 4) * LBO);
    const int Cp = pad_to(C, 128);
    float *bias_a1_s = wa0_s + kh * 12 + Cp;  // kernel_b_base: [wa0 | bias_ff (unused) | bias_a1]
    float *small_end = bias_a1_s + Cp;
    Ring ring;
    ring.slots = smem + pad_to((int)(reinterpret_cast<uint8_t *>(small_end) - smem), 128);
    ring.nslots = p.ring_slots;
    ring.full = bars;
    ring.empty = bars + kMaxRing;
    ring.reset();
    uint64_t *bar_mma = bars + 2 * kMaxRing;

    SliceSeq prod;
    prod.n = 0;
    prod.chunk_major = 1;
    prod.st[prod.n++] = &p.a1;
    prod.reset();
    const int per_tile = prod.per_tile();
    const bool sticky = p.ring_sticky != 0;
    if (sticky) {
        SliceSeq tmp = prod;
        uint32_t o = 0;
        for (int i = 0; i < per_tile && i < kMaxRing; i++) {
            long long off;
            uint32_t bytes;
            tmp.next(off, bytes);
            ring.off[i] = o;
            o += bytes;
        }
    } else {
        for (int i = 0; i < kMaxRing; i++) ring.off[i] = (uint32_t)i * kSlotBytes;
    }
    // One TMEM region per channel chunk (<= 4 x 128 columns = all of TMEM, one CTA per SM): the MMAs of EVERY
    // chunk of a tile are issued up-front, back to back, and the epilogue of chunk j runs while the tensor
    // core works on chunks j+1..  (the per-chunk lock step left the SM idle two thirds of the time).
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, (uint32_t)(Cp / 128 <= 1 ? 128 : (Cp / 128 == 2 ? 256 : 512)));
    if (tid == 0) {
        for (int i = 0; i < ring.nslots; i++) {
            tc::mbar_init(&ring.full[i], 1);
            tc::mbar_init(&ring.empty[i], 1);
        }
        for (int j = 0; j < kWideChunks; j++) tc::mbar_init(&bar_mma[j], 1);
        tc::mbar_init_fence();
    }
    for (int i = tid; i < kh * 8; i += kWideThreads)  // folded rows, see att0_folded_weight
        wa0_s[i] = att0_folded_weight(p.a0_w, p.a0_b, p.a0_cin, p.a0_cout, i >> 3, i & 7);
    for (int i = tid; i < Cp; i += kWideThreads) bias_a1_s[i] = i < C ? __ldg(p.a1.bias + i) : 0.f;
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total_slices = (long long)my_tiles * per_tile;
    if (warp == 8 && lane == 0) {
        ring_start(ring, prod, seq_off_s, seq_bytes_s, p.packed, per_tile, total_slices, sticky);
    }
    uint32_t mma_phase = 0;
    const long long rows_total = (long long)c.B * c.Nprev;
    const long long centers_total = (long long)c.B * c.O;
    const int row_w = 4 + c.Cin;
    const int out_w = 4 + C;
    const int attfdim = c.attfdim, Nprev = c.Nprev, O = c.O;
    const uint32_t row_off = (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
    const int nchunk = Cp / 128;
    const float pre_floor = c.pre_relu ? 0.f : -3.402823466e+38f;

    const int my_cl = row / K, my_slot = row - my_cl * K;
    bool pf_valid = false;
    int pf_idx = 0;
    long long pf_center = 0;
    uint32_t pf_roff = 0;
    float4 pf_head = make_float4(0.f, 0.f, 0.f, 0.f), pf_cent = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch_row = [&]() {
        pf_roff = 0;
        if (pf_valid) {
            const int b = (int)(pf_center / O);
            const long long r = take_row(pf_idx, b, Nprev, rows_total);
            pf_roff = (uint32_t)r * (uint32_t)C;
            const float *src = c.table + r * row_w;
            if ((row_w & 3) == 0) pf_head = __ldg(reinterpret_cast<const float4 *>(src));
            else pf_head = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 0.f);
            pf_cent = __ldg(c.cent + pf_center);
        }
    };
    if (warp < 8 && blockIdx.x < num_tiles) {
        const long long center = (long long)blockIdx.x * cpt + my_cl;
        pf_valid = my_cl < cpt && center < centers_total;
        pf_center = center;
        if (pf_valid) pf_idx = __ldg(c.nebidx + center * K + my_slot);
        fetch_row();
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long c_base = (long long)tile * cpt;
        // ---- gather + attention stage 0: two threads per edge row, each half of the channel groups ----
        if (warp < 8) {
            float att[12];
#pragma unroll
            for (int i = 0; i < 12; i++) att[i] = 0.f;
            att[10] = 1.f;
            float dx, dy, dz;
            const uint32_t roff = pf_roff;
            if (pf_valid) att_vector(attfdim, pf_cent, pf_head.x, pf_head.y, pf_head.z, att, dx, dy, dz);
            const Att7 a7 = fold_att(attfdim, att);
            {
                const long long ncenter = (long long)(tile + gridDim.x) * cpt + my_cl;
                pf_valid = (tile + (int)gridDim.x) < num_tiles && my_cl < cpt && ncenter < centers_total;
                pf_center = ncenter;
                if (pf_valid) pf_idx = __ldg(c.nebidx + ncenter * K + my_slot);
            }
            if (half == 0) rowoff_s[row] = roff;
            const int gh = kh / 8;  // groups of 4 channels per half
            for (int g = half * gh; g < (half + 1) * gh; g++) {
                const float4 *w4 = reinterpret_cast<const float4 *>(wa0_s + g * 32);
                float hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float acc = att0_channel(w4[i * 2], w4[i * 2 + 1], a7);
                    tc::split_op<NSPLIT>(fmaxf(acc, 0.f), hi[i], lo[i]);
                }
                const uint32_t off = row_off + (uint32_t)g * LBO;
                *reinterpret_cast<float4 *>(xa_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                if (NSPLIT == 3)
                    *reinterpret_cast<float4 *>(xa_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();

        if (warp == 8) {  // every chunk's MMAs, back to back; chunk j signals bar_mma[j]
            if (lane == 0) {
                for (int j = 0; j < nchunk; j++) {
                    run_transposed_stage<NSPLIT>(p.a1.Kp, 1, ring, sticky, tc::smem_u32(xa_hi),
                                                 tc::smem_u32(xa_lo), LBO, 128, tmem + (uint32_t)(j * 128), 0);
                    tc::mma_commit(&bar_mma[j]);
                }
            }
            __syncwarp();
        }
        for (int j = 0; j < nchunk; j++) {
            // ---- epilogue: thread = channel (TMEM lane of quadrant warp & 3), columns [64 * half, +64) ----
            const int ch = j * 128 + (warp & 3) * 32 + lane;
            const bool chv = ch < C;
            const bool warp_on = warp < 8 && (j * 128 + (warp & 3) * 32) < C;
            const int col0 = 64 * half;
            const float *fbase = p.ftab + (chv ? ch : 0);
            // All 64 feature values of this thread's columns are requested before the wait on the chunk's MMAs
            // (they only depend on the row offsets): 64 loads in flight per thread instead of 16 -- with eight
            // warps per SM the L2 latency of these gathers is what the epilogue waits for.
            uint32_t gA[16], gB[16], f0[16], f1[16], f2[16], f3[16];
            auto gather16 = [&](uint32_t (&f)[16], int e0) {
#pragma unroll
                for (int i = 0; i < 16; i++) f[i] = __float_as_uint(__ldg(fbase + rowoff_s[e0 + i]));
            };
            if (warp_on) {
                gather16(f0, col0);
                gather16(f1, col0 + 16);
                gather16(f2, col0 + 32);
                gather16(f3, col0 + 48);
            }
            if (j == 0 && warp < 8) fetch_row();
            if (warp < 8) wait_bar(&bar_mma[j], mma_phase);
            tc::fence_after_sync();
            if (warp_on) {
                const float ba = bias_a1_s[chv ? ch : 0];
                const uint32_t lane_addr = tmem + (uint32_t)(j * 128) + ((uint32_t)((warp & 3) * 32) << 16);
                float *out_ch = c.out + 4 + ch;
                float m = -3.402823466e+38f;
                int pos = 0, cl = col0 / K;
                auto reduce16 = [&](const uint32_t (&g)[16], const uint32_t (&f)[16]) {
                    float pr[16];
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        pr[i] = __uint_as_float(f[i]) * fmaxf(__uint_as_float(g[i]) + ba, 0.f);  // :167 att * feats
                    if ((K & 15) == 0) {
                        float mm = pr[0];
#pragma unroll
                        for (int i = 1; i < 16; i++) mm = fmaxf(mm, pr[i]);
                        m = fmaxf(m, mm);
                        pos += 16;
                        if (pos == K) {
                            const long long center = c_base + cl;
                            if (chv && cl < cpt && center < centers_total)
                                out_ch[center * out_w] = fmaxf(m, pre_floor) * __ldg(c.centmsk + center);
                            pos = 0;
                            cl++;
                            m = -3.402823466e+38f;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            m = fmaxf(m, pr[i]);
                            if (++pos == K) {
                                const long long center = c_base + cl;
                                if (chv && cl < cpt && center < centers_total)
                                    out_ch[center * out_w] = fmaxf(m, pre_floor) * __ldg(c.centmsk + center);
                                pos = 0;
                                cl++;
                                m = -3.402823466e+38f;
                            }
                        }
                    }
                };
                tc::tmem_ld16(lane_addr + col0, gA);
                tc::tmem_ld_wait();
                tc::tmem_ld16(lane_addr + col0 + 16, gB);
                reduce16(gA, f0);
                tc::tmem_ld_wait();
                tc::tmem_ld16(lane_addr + col0 + 32, gA);
                reduce16(gB, f1);
                tc::tmem_ld_wait();
                tc::tmem_ld16(lane_addr + col0 + 48, gB);
                reduce16(gA, f2);
                tc::tmem_ld_wait();
                reduce16(gB, f3);
            }
        }
        mma_phase ^= 1;  // every chunk barrier completed once for this tile
        if (warp < 8) {
            for (int i = tid; i < cpt * 4; i += 256) {
                const long long center = c_base + i / 4;
                if (center < centers_total)
                    c.out[center * out_w + (i & 3)] = __ldg(reinterpret_cast<const float *>(c.cent + center) + (i & 3));
            }
        }
        // the next tile rewrites the operand image and the accumulators: every epilogue read must be done
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)(Cp / 128 <= 1 ? 128 : (Cp / 128 == 2 ? 256 : 512)));
}

// ------------------------------------------------------------------------------------------------
// Host side: stage tables, packing, launches.
// ------------------------------------------------------------------------------------------------
static void make_stage(TcStage &s, int transposed, int cin, int cout, const float *bias, long long &off) {
    s.transposed = transposed;
    s.Cin = cin;
    s.Cout = cout;
    s.Kp = pad_to(cin, 8);
    s.Np = pad_to(cout, transposed ? 128 : 16);
    s.w_off = off;
    s.bias = bias;
    off += 2LL * s.Np * s.Kp;
}

// Fills the stage tables of both kernels; returns the packed size in floats (or < 0 if unsupported).
// `src` receives, for every packed stage, the index of its weight matrix in ConvParams::w.
struct PackList {
    TcStage st[2 * GRIDGCN_MAX_STAGES];
    int widx[2 * GRIDGCN_MAX_STAGES];
    int n;
};

static long long build_tc_plan(const ConvParams &c, TcParams &p, PackList *pl) {
    p.c = c;
    long long off = 0;
    const int nf = c.n_feat;
    p.na = p.nfh = p.has_ff = p.f0_cuda = 0;
    p.has_att = c.attfdim > 0;
    if (pl) pl->n = 0;
    auto add = [&](TcStage &st, int transposed, int widx) {
        make_stage(st, transposed, c.cin[widx], c.cout[widx], c.bias[widx], off);
        if (pl) {
            pl->st[pl->n] = st;
            pl->widx[pl->n++] = widx;
        }
    };
    if (nf > 3) return -1;  // the slice sequence holds up to 3 streamed stages
    if (c.Cin > 0) {
        for (int s = 0; s < nf; s++) add(p.a[p.na++], 1, s);
    } else {
        int first = 0;
        if (nf >= 2) {  // feature stage 0 (K = 3) runs on the CUDA cores
            p.f0_cuda = 1;
            p.f0_cout = c.cout[0];
            p.f0_w = c.w[0];
            p.f0_b = c.bias[0];
            first = 1;
        }
        for (int s = first; s + 1 < nf; s++) add(p.fh[p.nfh++], 0, s);
        add(p.ff, 1, nf - 1);
        p.has_ff = 1;
    }
    if (p.has_att) {  // attention stage 0 (K <= 10) runs on the CUDA cores
        p.a0_cin = c.cin[nf];
        p.a0_cout = c.cout[nf];
        p.a0_w = c.w[nf];
        p.a0_b = c.bias[nf];
        if (p.a0_cin > 10) return -1;
        add(p.a1, 1, nf + 1);
    }
    for (int s = 0; s < p.nfh; s++)
        if (p.fh[s].Np > 128) return -1;
    p.tmem_cols = p.has_ff ? 256 : 128;
    return off;
}

static size_t kernel_a_smem(const TcParams &p, int TR, int nsplit, int slots) {
    int kmax = 0;
    for (int s = 0; s < p.na; s++) kmax = max(kmax, p.a[s].Kp);
    size_t x_img = (size_t)(kmax / 4) * (TR * 16 + 16);
    return pad_to((int)((nsplit == 3 ? 2 : 1) * x_img), 128) + (size_t)slots * kSlotBytes + 1024;
}

// Shared-memory bytes of kernel B in front of the ring, and the per-tile slice sequence of the ring.
static size_t kernel_b_base(const TcParams &p, int nsplit) {
    int kx = 0;
    if (p.f0_cuda) kx = pad_to(p.f0_cout, 8);
    for (int s = 0; s < p.nfh; s++) kx = max(kx, max(p.fh[s].Kp, pad_to(p.fh[s].Cout, 8)));
    if (p.has_ff) kx = max(kx, p.ff.Kp);
    int kh = p.has_att ? p.a1.Kp : 0;
    size_t nimg = nsplit == 3 ? 2 : 1;
    size_t bytes = nimg * (size_t)(kx / 4 + kh / 4) * kTileRows * 16;
    for (int s = 0; s < p.nfh; s++) bytes += (size_t)2 * p.fh[s].Np * p.fh[s].Kp * 4;
    bytes += (size_t)kh * 12 * 4;                                   // wa0_s[kh][12]
    if (p.f0_cuda) bytes += (size_t)pad_to(p.f0_cout, 8) * 4 * 4;     // wf0_s[kf0][4]
    bytes += (size_t)(p.nfh * 128 + 2 * pad_to(p.c.Cout, 128)) * 4;    // biases
    return pad_to((int)bytes, 128);
}

static void kernel_b_sequence(const TcParams &p, int &n_slices, size_t &seq_bytes) {
    n_slices = 0;
    seq_bytes = 0;
    const TcStage *st[2];
    int n = 0;
    if (p.has_ff) st[n++] = &p.ff;
    if (p.has_att) st[n++] = &p.a1;
    for (int i = 0; i < n; i++) {
        n_slices += (st[i]->Np / 128) * ((st[i]->Kp + kSliceK - 1) / kSliceK);
        seq_bytes += (size_t)2 * st[i]->Np * st[i]->Kp * 4;
    }
}

int launch_first_ws(const TcParams &p, cudaStream_t st);  // gridconv_first_ws.cu; -1: layer does not fit
int launch_edge_ws(const TcParams &p, cudaStream_t st);   // gridconv_edge_ws.cu;  -1: layer does not fit
int launch_rowgemm_tc(const float *in1, int ld1, int c1, const float *in2, int ld2, int c2, const float *W,
                      const float *bias, int N, int relu_in, int relu_out, const float *scale, float *out, int ldo,
                      const float *cent, float *out_table, long long rows, cudaStream_t st);  // rowgemm_tc.cu

// Kernel A as a chain of row GEMMs (r02): every stage of the per-point feature MLP is one launch of the persistent
// tensor-core GEMM of rowgemm_tc.cu (stages wider than 256 outputs: one launch per 256 columns); the hidden
// activations ping-pong through two workspace buffers behind F.  Replaces the fused, lock-step
// point_mlp_*_tc_kernel (13-17 % of the tensor pipe, r01 profile) whenever the views are 16-byte aligned.
static int launch_point_mlp_chain(const TcParams &p, cudaStream_t st) {
    const ConvParams &c = p.c;
    if (p.nsplit != 3 || c.Cin <= 0 || (c.Cin & 3)) return -1;
    for (int s = 0; s < c.n_feat; s++)
        if (c.cout[s] & 3) return -1;
    const long long rows = (long long)c.B * c.Nprev;
    // GRIDGCN_POINT_MLP_CHAIN_MIN_ROWS: below this many source points the fused kernel-A variants run instead
    static const long long min_rows = [] { const char *e = getenv("GRIDGCN_POINT_MLP_CHAIN_MIN_ROWS"); return e ? atoll(e) : 0LL; }();
    if (rows < min_rows) return -1;
    int hmax = 0;
    for (int s = 0; s + 1 < c.n_feat; s++) hmax = std::max(hmax, c.cout[s]);
    float *tmp[2] = {p.ftab + rows * c.Cout, p.ftab + rows * ((long long)c.Cout + hmax)};
    const float *src = c.table + 4;
    int ld = 4 + c.Cin, cw = c.Cin;
    for (int s = 0; s < c.n_feat; s++) {
        float *dst = s + 1 == c.n_feat ? p.ftab : tmp[s & 1];
        const int N = c.cout[s];
        for (int n0 = 0; n0 < N; n0 += 256) {
            const int rc = launch_rowgemm_tc(src, ld, cw, nullptr, 0, 0, c.w[s] + (size_t)n0 * cw, c.bias[s] + n0,
                                             std::min(256, N - n0), 0, 1, nullptr, dst + n0, N, nullptr, nullptr, rows, st);
            if (rc != 0) return rc < 0 ? (s == 0 && n0 == 0 ? -1 : GRIDGCN_ELIMIT) : rc;
        }
        src = dst;
        ld = cw = N;
    }
    return 0;
}

constexpr size_t kSmemCap = 224 * 1024;  // dynamic part; static barriers/index cache use < 3 KB of the 227 KB

static void launch_first64(const TcParams &p, int blocks, size_t smem, cudaStream_t st, int tiles, int cpt,
                           int hid, int stash_lo) {
    int dev = 0;
    cudaGetDevice(&dev);
    static PerDeviceOnce attr_set;
    if (!attr_set.done(dev)) {
        cudaFuncSetAttribute(edge_first64_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap);
        attr_set.set(dev);
    }
    edge_first64_kernel<3><<<blocks, kTcThreads, smem, st>>>(p, tiles, cpt, hid, stash_lo);
}

template <int NSPLIT>
static int launch_tc_t(TcParams &p, cudaStream_t st) {
    const ConvParams &c = p.c;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static PerDeviceOnce attr_set;
    if (!attr_set.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(point_mlp_tc_kernel<NSPLIT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(point_mlp_rows_tc_kernel<NSPLIT>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(edge_tc_kernel<NSPLIT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kSmemCap);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(edge_tc_kernel<NSPLIT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kSmemCap);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(edge_wide_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kSmemCap);
        if (e != cudaSuccess) return (int)e;
        attr_set.set(dev);
    }
    int chain_rc = -1;
    if (p.na > 0 && NSPLIT == 3) {
        chain_rc = launch_point_mlp_chain(p, st);
        if (chain_rc > 0) return chain_rc;
    }
    if (p.na > 0 && chain_rc != 0) {  // kernel A, fused variants
        int chunks = 0, kmax = 0, npmax = 0;
        for (int s = 0; s < p.na; s++) {
            chunks = max(chunks, p.a[s].Np / 128);
            kmax = max(kmax, p.a[s].Kp);
            npmax = max(npmax, p.a[s].Np);
        }
        long long rows = (long long)c.B * c.Nprev;
        // Row-major variant when a [128 x kmax] hi/lo image plus >= 2 ring slots fit and C % 4 == 0
        const size_t rows_fixed = (size_t)(NSPLIT == 3 ? 2 : 1) * (kmax / 4) * kTileRows * 16 +
                                  pad_to(p.na * npmax * 4, 128) + 1024;
        const bool rows_ok = rows_fixed + 2 * (size_t)kSlotBytes <= kSmemCap && npmax <= 512 &&
                             (p.a[p.na - 1].Cout & 3) == 0;
        if (rows_ok) {
            p.a_rows = kTileRows;
            p.ring_slots = (int)min((size_t)GG_RING_CAP, (kSmemCap - rows_fixed) / kSlotBytes);
            p.ring_sticky = 0;
            int saved_cols = p.tmem_cols;
            p.tmem_cols = npmax <= 128 ? 128 : (npmax <= 256 ? 256 : 512);
            long long tiles = (rows + kTileRows - 1) / kTileRows;
            size_t smem = rows_fixed + (size_t)p.ring_slots * kSlotBytes;
            int per_sm = (int)max((size_t)1, min(min((size_t)2, kSmemCap / smem), (size_t)(512 / p.tmem_cols)));
            int blocks = (int)min(tiles, (long long)sms * per_sm);
            point_mlp_rows_tc_kernel<NSPLIT><<<blocks, kRowsThreads, smem, st>>>(p, (int)tiles);
            p.tmem_cols = saved_cols;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return (int)e;
        } else {
            // Transposed variant.  Rows per tile: the largest of 128/64/32 that fits chunks*TR <= 256 TMEM
            // columns and leaves at least 2 ring slots (up to 6 when there is room).  Fewer, larger tiles
            // win: every tile re-streams the layer's weights from L2 and that stream -- ~2-2.5 TB/s in
            // aggregate when all SMs read the same slices -- is the bound for the wide layers (measured, r01).
            int TR = 0, slots = 0;
            for (int cand = 128; cand >= 32; cand >>= 1) {
                if (chunks * cand > 256) continue;
                size_t fixed = kernel_a_smem(p, cand, NSPLIT, 0);
                if (fixed > kSmemCap) continue;
                int fit = (int)min((size_t)min(GG_RING_CAP + 2, kMaxRing), (kSmemCap - fixed) / kSlotBytes);
                if (fit >= 2) { TR = cand; slots = fit; break; }
            }
            if (TR == 0) return GRIDGCN_ELIMIT;
            p.a_rows = TR;
            p.ring_slots = slots;
            p.ring_sticky = 0;
            long long tiles = (rows + TR - 1) / TR;
            size_t smem = kernel_a_smem(p, TR, NSPLIT, slots);
            int per_sm = (int)max((size_t)1, min((size_t)2, kSmemCap / smem));
            int blocks = (int)min(tiles, (long long)sms * per_sm);
            point_mlp_tc_kernel<NSPLIT><<<blocks, kRowsThreads, smem, st>>>(p, (int)tiles);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return (int)e;
        }
    }
    {  // kernel B
        if (c.K > kTileRows) return GRIDGCN_ELIMIT;
        if ((long long)c.B * c.Nprev * c.Cout >= (1LL << 32)) return GRIDGCN_ELIMIT;  // 32-bit row offsets
        if (NSPLIT == 3 && p.has_ff) {  // persistent warp-specialised first-layer pipeline (gridconv_first_ws.cu)
            const int rc = launch_first_ws(p, st);
            if (rc >= 0) return rc;
        }
        if (NSPLIT == 3 && !p.has_ff) {  // weight-stationary pipeline for the layers with input features (gridconv_edge_ws.cu)
            const int rc = launch_edge_ws(p, st);
            if (rc >= 0) return rc;
        }
        if (NSPLIT == 3 && p.has_ff && p.has_att && p.f0_cuda && p.dbg == nullptr && c.Cout <= 64 &&
            (64 % c.K == 0 || c.K == 128)) {
            // compact first-layer variant: three CTAs per SM (see edge_first64_kernel)
            const First64Layout L = first64_layout(p);
            int hid = 32;
            for (int s = 0; s < p.nfh; s++)
                while (hid < p.fh[s].Np) hid <<= 1;
            const size_t smem = L.total + 128;
            const int per_sm = (int)min(min((size_t)3, kSmemCap / smem), (size_t)(512 / (128 + hid)));
            if (2 * L.kh <= L.kx && p.ff.Np == 128 && p.a1.Np == 128 && per_sm >= 3) {
                const int cpt = kTileRows / c.K;
                const long long centers = (long long)c.B * c.O;
                const long long tiles = (centers + cpt - 1) / cpt;
                if (tiles > 0x7fffffff) return GRIDGCN_ELIMIT;
                const int blocks = (int)min(tiles, (long long)sms * per_sm);
                launch_first64(p, blocks, smem, st, (int)tiles, cpt, hid, L.kf0 <= hid ? 1 : 0);
                return (int)cudaGetLastError();
            }
        }
        const size_t base = kernel_b_base(p, NSPLIT);
        int n_slices;
        size_t seq_bytes;
        kernel_b_sequence(p, n_slices, seq_bytes);
        size_t smem;
        if (n_slices <= kMaxRing && base + seq_bytes + 1024 <= kSmemCap) {
            p.ring_sticky = 1;  // every weight slice of a tile stays resident: loaded once per CTA
            p.ring_slots = max(n_slices, 1);
            smem = base + seq_bytes + 1024;
        } else {
            p.ring_sticky = 0;
            if (base + 1024 > kSmemCap) return GRIDGCN_ELIMIT;
            p.ring_slots = (int)min((size_t)GG_RING_CAP, (kSmemCap - base - 1024) / kSlotBytes);
            if (p.ring_slots < 2) return GRIDGCN_ELIMIT;
            smem = base + (size_t)p.ring_slots * kSlotBytes + 1024;
        }
        const int cpt = kTileRows / c.K;
        long long centers = (long long)c.B * c.O;
        long long tiles = (centers + cpt - 1) / cpt;
        if (tiles > 0x7fffffff) return GRIDGCN_ELIMIT;
        // co-resident CTAs hide the lock-step latencies: limited by shared memory, TMEM columns (512 per SM)
        // and threads
        int per_sm = (int)min(min((size_t)(p.has_ff ? 2 : 3), kSmemCap / smem), (size_t)(512 / p.tmem_cols));  // 3: register limit (128 regs x 160 threads)
        per_sm = max(per_sm, 1);
        int blocks = (int)min(tiles, (long long)sms * per_sm);
        if (p.has_ff)
            edge_tc_kernel<NSPLIT, true><<<blocks, kTcThreads, smem, st>>>(p, (int)tiles, cpt);
        else if ((per_sm == 1 || GG_WIDE_ALWAYS) && p.has_att && p.dbg == nullptr && 64 % c.K == 0 &&
                 (p.a1.Kp & 7) == 0 && n_slices <= 32 &&
                 pad_to(c.Cout, 128) / 128 <= 4) {  // slice table / one TMEM region per chunk
            // 288 threads x ~160 registers: one CTA per SM whatever shared memory would allow
            edge_wide_kernel<NSPLIT><<<(int)min(tiles, (long long)sms), kWideThreads, smem, st>>>(p, (int)tiles, cpt);
        }
        else
            edge_tc_kernel<NSPLIT, false><<<blocks, kTcThreads, smem, st>>>(p, (int)tiles, cpt);
        return (int)cudaGetLastError();
    }
}

int tc_packed_floats(const ConvParams &c) {
    TcParams p{};
    long long n = build_tc_plan(c, p, nullptr);
    return n < 0 || n > 0x7fffffff ? -1 : (int)max(n, 4LL);
}

int tc_pack(const ConvParams &c, float *packed, cudaStream_t st) {
    TcParams p{};
    PackList pl;
    if (build_tc_plan(c, p, &pl) < 0) return GRIDGCN_ELIMIT;
    for (int i = 0; i < pl.n; i++) {
        const TcStage &s = pl.st[i];
        int total = s.Np * s.Kp;
        pack_stage_kernel<<<(total + 255) / 256, 256, 0, st>>>(c.w[pl.widx[i]], packed, s);
    }
    return (int)cudaGetLastError();
}

static unsigned long long *g_phase_buf = nullptr;
void tc_set_phase_buffer(unsigned long long *buf) { g_phase_buf = buf; }

int launch_gridconv_tc(const ConvParams &c, int precision, const float *packed, float *ftab,
                       cudaStream_t st) {
    TcParams p{};
    if (build_tc_plan(c, p, nullptr) < 0) return GRIDGCN_ELIMIT;
    if (!packed || (c.Cin > 0 && !ftab)) return GRIDGCN_EWORKSPACE;
    p.packed = packed;
    p.ftab = ftab;
    p.dbg = g_phase_buf;
    p.nsplit = precision == GRIDGCN_PRECISION_TF32X3 ? 3 : 1;
    return p.nsplit == 3 ? launch_tc_t<3>(p, st) : launch_tc_t<1>(p, st);
}

}  // namespace gg
