// GridConv layers WITH input features (layers >= 1 of the encoder, decoder cores) as a persistent, warp-specialised,
// WEIGHT-STATIONARY tcgen05 pipeline -- GRIDGCN_PRECISION_TF32X3.  Replaces edge_tc_kernel<3,false> /
// edge_wide_kernel (gridconv_tc.cu) where its shapes fit.
//
// Reference semantics: segmentation/models/gcn_module_g_att.py:120-170 (attention MLP, product with the features),
// :45-79 (max pool over the K slots), :172-287, gathers of utils/ops.py:78-93.  The per-point feature MLP has been
// hoisted (gridconv_tc.cu: F = MLP(table) once per source point); what remains per (centre, neighbour) edge is
//     att = relu(W1 relu(W0 att_vec + b0) + b1),   out[centre, ch] = max_k att[k, ch] * F[nbr(k), ch].
//
// r01/r02 profiles of the kernels this replaces: lock-step phases (tensor pipe 27 %, 14 % warps active) and, for the
// 512-channel layer, 64 % of the time waiting for the attention weights that every 128-edge tile re-streams (0.5 MB per
// tile).  Here every CTA owns ONE 128-channel chunk of the attention output for the whole launch and keeps that
// chunk's weights in TENSOR MEMORY (A operand of a transposed M = 128 MMA: no operand fetch from shared memory at all);
// the units of 128 edges stream through
//     G   warps 0-3    gather: neighbour index -> table row (xyz) + centre -> input image X0 [128 x 8], F-row offsets
//     MS  warp 16      stage 0 on the tensor core: D0[edge, A0p] = X0 * W0^T   (folded attention stage 0, K = 8)
//     E0  warps 4-7    D0 -> relu -> hi/lo operand image xa [128 x A0p] (shared memory, B operand)
//     MA  warp 17      D^T[ch, edge] (+)= Wchunk[tmem] * xa^T,  A0p/8 k-steps x 3 passes, N = 128 edges
//     EF  warps 8-15   thread = channel x 64-edge half: F gather (coalesced over channels), relu(att + b) * f, running
//                      max over each centre's K consecutive columns, store
// with mbarrier rings between the roles (one arrival per warp; tcgen05.commit from the MMA threads).  Stage 0 and its
// epilogue are recomputed by each of the C/128 chunk owners of a unit -- a few percent of the work, against streaming
// 2 * A0p * C * 4 bytes of weights per unit.
//
// TMEM (512 columns): weights hi | lo (2 * A0p), stage-0 accumulator ring, D^T ring:
//     A0p = 32:  64 + 2 x 32 + 3 x 128      A0p = 64:  128 + 2 x 64 + 2 x 128      A0p = 128:  256 + 2 x 128 (S0 and D^T of a unit share a slot)
#include "gridconv_tc.cuh"

#include <algorithm>
#include <cstdlib>

namespace gg {

constexpr int kEwThreads = 576;   // 18 warps
constexpr int kEwRo = 8;          // ring of F-row-offset tables (G runs ahead of EF by at most the other rings' depths)
constexpr uint32_t kEwPanel = 2048;

template <int D>
struct EwRing {
    int slot = 0;
    uint32_t ph = 0;
    __device__ __forceinline__ void next() {
        if (++slot == D) { slot = 0; ph ^= 1u; }
    }
};

__device__ __forceinline__ void ew_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void ew_tmem_st8(uint32_t taddr, const float *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

template <int A0P>
struct EwCfg {
    // A0P = 128: the weights take 256 of the 512 columns; the stage-0 accumulator of a unit and its D^T tile then SHARE a
    // 128-column slot (S0(u) is dead once E0 has turned ALL of it into the xa image -- MA(u) waits for both K halves
    // before its first MMA overwrites the slot), which gives both rings depth 2: EF drains D^T(u) while the tensor core
    // works on unit u + 1.
    static constexpr bool kWide = A0P > 64;
    static constexpr int DA = 2;                                   // stage-0 accumulator ring
    static constexpr int DI = A0P <= 32 ? 3 : (A0P <= 64 ? 2 : 1); // xa image ring
    static constexpr int DD = A0P <= 32 ? 3 : 2;                   // D^T ring
    static constexpr uint32_t kWCol = 0, kAccCol = 2 * A0P, kAccStride = kWide ? 128 : A0P;
    static constexpr uint32_t kDCol = kWide ? kAccCol : 2 * A0P + DA * A0P;
    static_assert(kDCol + DD * 128 <= 512, "TMEM map");
    static constexpr int KH = kWide ? 2 : 1;   // the (single) xa image of A0P = 128 is handed over in two K halves: MA multiplies
                                               // half 0 while E0 still writes half 1, and E0 starts the next unit's half 0
                                               // as soon as MA has read this unit's (E0 and MA took ~3000 cycles EACH, in turn)
    static constexpr uint32_t kImgBytes = 2u * (A0P / 4) * kEwPanel;  // hi | lo
    // shared memory: W0 hi | lo, bias chunk, X0 ring, row-offset ring, xa ring
    static constexpr uint32_t w0_hi = 0, w0_lo = A0P * 32, bias = 2 * A0P * 32, x0 = (bias + 512 + 127) & ~127u;
    static constexpr uint32_t ro = x0 + 2 * 4 * kEwPanel, img = ro + kEwRo * 512, total = img + DI * kImgBytes;
};

template <int A0P>
__global__ void __launch_bounds__(kEwThreads, 1)
edge_ws_kernel(const __grid_constant__ TcParams p, int num_units, int cpt, int log2k, int nchunk, int blocked) {
    using Cf = EwCfg<A0P>;
    constexpr int DA = Cf::DA, DI = Cf::DI, DD = Cf::DD;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[4 + 2 * DA + 2 * DI * Cf::KH + 2 * DD + 2 * kEwRo];
    __shared__ uint32_t tmem_base_s;
    __shared__ float pair_max[4][32];  // K = 128: the second half's partial maxima
    uint64_t *x0_full = bars, *x0_free = bars + 2;             // G  -> MS (4); MS -> G (commit: the stage-0 MMAs have read X0)
    uint64_t *s0_done = x0_free + 2, *acc_free = s0_done + DA;  // MS -> E0 (commit); E0 -> MS (4)
    constexpr int KH = Cf::KH;
    uint64_t *e0_done = acc_free + DA, *img_free = e0_done + DI * KH;  // E0 -> MA (4); MA -> E0 (commit); per K half
    uint64_t *d_full = img_free + DI * KH, *d_free = d_full + DD;    // MA -> EF (commit); EF -> MA (8)
    uint64_t *ro_free = d_free + DD;                            // EF -> G (8)
    uint64_t *ro_full = ro_free + kEwRo;                        // G -> EF (4): row offsets of a unit are written
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ConvParams &c = p.c;
    const int C = c.Cout;
    const int chunk = (int)blockIdx.x % nchunk;                 // this CTA's 128 output channels
    const int owners = (int)gridDim.x / nchunk, me = (int)blockIdx.x / nchunk;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 32) {
        for (int i = 0; i < 2; i++) { tc::mbar_init(&x0_full[i], 4); tc::mbar_init(&x0_free[i], 1); }
        for (int i = 0; i < DA; i++) { tc::mbar_init(&s0_done[i], 1); tc::mbar_init(&acc_free[i], 4); }
        for (int i = 0; i < DI * KH; i++) { tc::mbar_init(&e0_done[i], 4); tc::mbar_init(&img_free[i], 1); }
        for (int i = 0; i < DD; i++) { tc::mbar_init(&d_full[i], 1); tc::mbar_init(&d_free[i], 8); }
        for (int i = 0; i < kEwRo; i++) { tc::mbar_init(&ro_free[i], 8); tc::mbar_init(&ro_full[i], 4); }
        tc::mbar_init_fence();
    }
    {   // folded attention stage 0 as the K = 8 image W0[A0P x 8] (att0_folded_weight), bias of this chunk
        const uint32_t lbo0 = (uint32_t)A0P * 16u;
        for (int i = tid; i < A0P * 8; i += kEwThreads) {
            const int n = i >> 3, k = i & 7;
            float hi, lo;
            tc::split_tf32(att0_folded_weight(p.a0_w, p.a0_b, p.a0_cin, p.a0_cout, n, k), hi, lo);
            const uint32_t off = tc::kmajor_off((uint32_t)n, (uint32_t)k, lbo0);
            *reinterpret_cast<float *>(smem + Cf::w0_hi + off) = hi;
            *reinterpret_cast<float *>(smem + Cf::w0_lo + off) = lo;
        }
        for (int i = tid; i < 128; i += kEwThreads) {
            const int ch = chunk * 128 + i;
            reinterpret_cast<float *>(smem + Cf::bias)[i] = ch < C ? __ldg(p.a1.bias + ch) : 0.f;
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t q = (uint32_t)(warp & 3);
    const uint32_t lane_base = (q * 32u) << 16;
    if (warp < 4) {  // this chunk's attention stage-1 weights -> TMEM (A operand: row = lane, k = column), hi | lo
        const int ch = chunk * 128 + (int)q * 32 + lane;
        const float *wrow = c.w[c.n_feat + 1] + (size_t)ch * p.a1.Cin;
#pragma unroll 1
        for (int k0 = 0; k0 < A0P; k0 += 8) {
            float hi[8], lo[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float w = (ch < C && k0 + k < p.a1.Cin) ? __ldg(wrow + k0 + k) : 0.f;
                tc::split_tf32(w, hi[k], lo[k]);
            }
            ew_tmem_st8(tmem + lane_base + Cf::kWCol + (uint32_t)k0, hi);
            ew_tmem_st8(tmem + lane_base + Cf::kWCol + (uint32_t)(A0P + k0), lo);
        }
        tc::tmem_st_wait();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    // units per owner: CONSECUTIVE units (blocked != 0: a CTA stays inside one cloud for many units, so the feature rows
    // its EF warps gather are re-used out of L1 -- the gathers were L2-bandwidth bound, r02) or round-robin
    const int u_base = num_units / owners, u_rem = num_units % owners;
    const int u_first = blocked ? me * u_base + min(me, u_rem) : me;
    const int u_step = blocked ? 1 : owners;
    const int n_my = blocked ? u_base + (me < u_rem ? 1 : 0) : (me < num_units ? (num_units - me + owners - 1) / owners : 0);
    const unsigned centers_total = (unsigned)c.B * (unsigned)c.O;
    const int out_w = 4 + C;
    auto unit_of = [&](int i) { return (unsigned)(u_first + i * u_step); };

    if (warp < 4) {
        // =========================== G: gather ===========================
        const int r = tid;
        const uint32_t row_off = ((uint32_t)r >> 3) * 128u + ((uint32_t)r & 7u) * 16u;
        const int my_cl = r >> log2k, my_slot = r & ((1 << log2k) - 1);
        const int Nprev = c.Nprev, O = c.O;
        const int rows_total = c.B * c.Nprev;
        const int row_w = 4 + c.Cin;
        auto center_of = [&](int i) -> unsigned { return unit_of(i) * (unsigned)cpt + (unsigned)my_cl; };
        auto load_idx = [&](int i, bool &valid) -> int {
            const unsigned center = center_of(i);
            valid = i < n_my && my_cl < cpt && center < centers_total;
            return valid ? __ldg(c.nebidx + ((size_t)center << log2k) + my_slot) : 0;
        };
        auto load_row = [&](int i, bool valid, int idx, float4 &head, float4 &cent, uint32_t &roff) {
            head = make_float4(0.f, 0.f, 0.f, 0.f);
            cent = head;
            roff = 0;
            if (valid) {
                const unsigned center = center_of(i);
                const int b = (int)(center / (unsigned)O);
                int gi = idx + b * Nprev;  // take(): clip after the batch offset (utils/ops.py:90-92)
                gi = gi < 0 ? 0 : (gi >= rows_total ? rows_total - 1 : gi);
                roff = (uint32_t)gi * (uint32_t)C;
                const float *src = c.table + (size_t)gi * row_w;
                head = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 0.f);
                cent = __ldg(c.cent + center);
            }
        };
        bool v_cur, v_nxt;
        int idx_nxt;
        float4 head, cent;
        uint32_t roff;
        {
            const int idx0 = load_idx(0, v_cur);
            load_row(0, v_cur, idx0, head, cent, roff);
            idx_nxt = load_idx(1, v_nxt);
        }
        EwRing<kEwRo> ro;
        for (int i = 0; i < n_my; i++) {
            const int x = i & 1;
            float4 head_n, cent_n;
            uint32_t roff_n;
            load_row(i + 1, v_nxt, idx_nxt, head_n, cent_n, roff_n);
            bool v_n2;
            const int idx_n2 = load_idx(i + 2, v_n2);
            float in[8];
            {
                const float dx = head.x - cent.x, dy = head.y - cent.y, dz = head.z - cent.z;
                const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
                in[0] = dist; in[1] = dx; in[2] = dy; in[3] = dz;
                in[4] = cent.x; in[5] = cent.y; in[6] = cent.z; in[7] = v_cur ? 1.f : 0.f;
            }
            float lo[8];
#pragma unroll
            for (int k = 0; k < 8; k++) lo[k] = in[k] - __uint_as_float(__float_as_uint(in[k]) & 0xFFFFE000u);
            if (i >= 2) tc::mbar_wait(&x0_free[x], (uint32_t)((i - 2) >> 1) & 1u);  // stage-0 MMAs of unit i-2 have read this slot
            if (i >= kEwRo) tc::mbar_wait(&ro_free[ro.slot], ro.ph ^ 1u);
            reinterpret_cast<uint32_t *>(smem + Cf::ro)[ro.slot * 128 + r] = roff;
            uint8_t *x0 = smem + Cf::x0 + (uint32_t)x * 4u * kEwPanel + row_off;
            *reinterpret_cast<float4 *>(x0) = make_float4(in[0], in[1], in[2], in[3]);
            *reinterpret_cast<float4 *>(x0 + kEwPanel) = make_float4(in[4], in[5], in[6], in[7]);
            *reinterpret_cast<float4 *>(x0 + 2 * kEwPanel) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4 *>(x0 + 3 * kEwPanel) = make_float4(lo[4], lo[5], lo[6], lo[7]);
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(&x0_full[x]);
                tc::mbar_arrive(&ro_full[ro.slot]);  // EF prefetches the feature rows of a unit ahead of its MMAs
            }
            if (chunk == 0 && r < cpt * 4) {  // centre columns of the output rows: written by the chunk-0 owner
                const unsigned center = unit_of(i) * (unsigned)cpt + (unsigned)(r >> 2);
                if (center < centers_total)
                    c.out[(size_t)center * out_w + (r & 3)] = __ldg(reinterpret_cast<const float *>(c.cent + center) + (r & 3));
            }
            head = head_n; cent = cent_n; roff = roff_n; v_cur = v_nxt;
            idx_nxt = idx_n2; v_nxt = v_n2;
            ro.next();
        }
    } else if (warp < 8) {
        // =========================== E0: stage-0 epilogue -> xa image ===========================
        const uint32_t row = q * 32u + (uint32_t)lane;
        const uint32_t row_off = (row >> 3) * 128u + (row & 7u) * 16u;
        EwRing<DA> a;
        EwRing<DI> d;
        for (int i = 0; i < n_my; i++) {
            tc::mbar_wait(&s0_done[a.slot], a.ph);
            tc::fence_after_sync();
            const uint32_t taddr = tmem + lane_base + Cf::kAccCol + (uint32_t)a.slot * Cf::kAccStride;
            uint8_t *img = smem + Cf::img + (uint32_t)d.slot * Cf::kImgBytes + row_off;
#pragma unroll 1
            for (int h = 0; h < KH; h++) {
                if (i >= DI) tc::mbar_wait(&img_free[d.slot * KH + h], d.ph ^ 1u);
#pragma unroll 1
                for (int c0 = h * (A0P / KH); c0 < (h + 1) * (A0P / KH); c0 += 16) {
                    uint32_t v[16];
                    tc::tmem_ld16(taddr + (uint32_t)c0, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        float x[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) x[j] = fmaxf(__uint_as_float(v[4 * g + j]), 0.f);
                        tc::split_lo2(x[0], x[1], lo[0], lo[1]);  // packed pairs: bit-identical to x - trunc(x)
                        tc::split_lo2(x[2], x[3], lo[2], lo[3]);
                        uint8_t *dst = img + (uint32_t)((c0 >> 2) + g) * kEwPanel;
                        *reinterpret_cast<float4 *>(dst) = make_float4(x[0], x[1], x[2], x[3]);
                        *reinterpret_cast<float4 *>(dst + (A0P / 4) * kEwPanel) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
                tc::fence_async_smem();
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    if (h == KH - 1) tc::mbar_arrive(&acc_free[a.slot]);
                    tc::mbar_arrive(&e0_done[d.slot * KH + h]);
                }
            }
            a.next();
            d.next();
        }
    } else if (warp < 16) {
        // =========================== EF: final epilogue ===========================
        const int hh = (warp - 8) >> 2;               // columns (edges) [64 hh, 64 hh + 64)
        const int chl = (int)q * 32 + lane;           // channel within the chunk == TMEM lane
        const int ch = chunk * 128 + chl;
        const bool chv = ch < C;
        const float ba = reinterpret_cast<const float *>(smem + Cf::bias)[chl];
        const float *fbase = p.ftab + (chv ? ch : 0);
        float *out_ch = c.out + 4 + ch;
        const int kmask = (1 << log2k) - 1;
        EwRing<DD> dd;
        EwRing<kEwRo> rc;  // ring position of the current unit's row offsets
        // The gathered features (this channel x the 64 edges of this half = four groups of 16 per unit) are prefetched TWO
        // GROUPS AHEAD through a 32-register window: when group k has been consumed its registers are re-loaded with group
        // k + 2 (of this unit, or groups 0 / 1 of the next one), so the L2 latency of the gather overlaps the compute of
        // the group in between and the wait for the next unit's MMAs (before: four exposed round trips per unit, the
        // kernel's bound -- r02 profile).
        float f[32];
        auto load_group = [&](const uint32_t *roff, int c0, int slot) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {  // coalesced over the warp (32 consecutive channels of one row)
                const uint4 o4 = *reinterpret_cast<const uint4 *>(roff + c0 + j);
                f[slot * 16 + j] = __ldg(fbase + o4.x); f[slot * 16 + j + 1] = __ldg(fbase + o4.y);
                f[slot * 16 + j + 2] = __ldg(fbase + o4.z); f[slot * 16 + j + 3] = __ldg(fbase + o4.w);
            }
        };
        if (n_my > 0) {
            tc::mbar_wait(&ro_full[0], 0u);
            const uint32_t *roff = reinterpret_cast<const uint32_t *>(smem + Cf::ro) + 64 * hh;
            load_group(roff, 0, 0);
            load_group(roff, 16, 1);
        }
        for (int i = 0; i < n_my; i++) {
            const unsigned c_base = unit_of(i) * (unsigned)cpt;
            const bool has_next = i + 1 < n_my;
            EwRing<kEwRo> rn = rc;
            rn.next();
            tc::mbar_wait(&d_full[dd.slot], dd.ph);
            tc::fence_after_sync();
            if (has_next) tc::mbar_wait(&ro_full[rn.slot], rn.ph);
            const uint32_t taddr = tmem + lane_base + Cf::kDCol + (uint32_t)dd.slot * 128u + (uint32_t)(64 * hh);
            const uint32_t *roff_c = reinterpret_cast<const uint32_t *>(smem + Cf::ro) + rc.slot * 128 + 64 * hh;
            const uint32_t *roff_n = reinterpret_cast<const uint32_t *>(smem + Cf::ro) + rn.slot * 128 + 64 * hh;
            float m = -3.402823466e+38f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c0 = 16 * k, slot = k & 1;
                uint32_t g[16];
                tc::tmem_ld16(taddr + (uint32_t)c0, g);
                tc::tmem_ld_wait();
                float mm = -3.402823466e+38f;
#pragma unroll
                for (int j = 0; j < 16; j += 2) {  // :167 att * feats; the features are ReLU outputs (>= 0), so
                    // f * relu(a) == max(f * a, 0): the attention ReLU is applied once per centre below.  Packed pairs.
                    float a0 = __uint_as_float(g[j]), a1 = __uint_as_float(g[j + 1]);
                    tc::add2(a0, a1, ba, ba);
                    tc::mul2(a0, a1, f[slot * 16 + j], f[slot * 16 + j + 1]);
                    mm = fmaxf(mm, fmaxf(a0, a1));
                }
                if (k < 2) load_group(roff_c, c0 + 32, slot);
                else if (has_next) load_group(roff_n, c0 - 32, slot);
                m = fmaxf(m, mm);
                const int e_end = 64 * hh + c0 + 16;
                if (log2k <= 6 && (e_end & kmask) == 0) {
                    const unsigned center = c_base + (unsigned)((e_end >> log2k) - 1);
                    if (chv && center < centers_total)
                        out_ch[(size_t)center * out_w] = fmaxf(m, 0.f) * __ldg(c.centmsk + center);
                    m = -3.402823466e+38f;
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(&d_free[dd.slot]);
                tc::mbar_arrive(&ro_free[rc.slot]);
            }
            rc = rn;
            if (log2k == 7) {  // one centre per unit: the two 64-edge halves meet through shared memory
                if (hh == 1) pair_max[q][lane] = m;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)q) : "memory");
                if (hh == 0) {
                    m = fmaxf(m, pair_max[q][lane]);
                    if (chv && c_base < centers_total)
                        out_ch[(size_t)c_base * out_w] = fmaxf(m, 0.f) * __ldg(c.centmsk + c_base);
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)q) : "memory");
            }
            dd.next();
        }
    } else if (lane == 0) {
        const uint32_t sb = tc::smem_u32(smem);
        auto adv = [](uint64_t dsc, uint32_t bytes) -> uint64_t { return dsc + (uint64_t)(bytes >> 4); };
        if (warp == 16) {
            // =========================== MS: stage 0 ===========================
            const uint32_t idesc = tc::make_idesc_tf32(128, A0P);
            const uint32_t lbo_w0 = (uint32_t)A0P * 16u;
            const uint64_t w0h = tc::make_sdesc(sb + Cf::w0_hi, lbo_w0), w0l = tc::make_sdesc(sb + Cf::w0_lo, lbo_w0);
            const uint64_t xh0 = tc::make_sdesc(sb + Cf::x0, kEwPanel), xl0 = tc::make_sdesc(sb + Cf::x0 + 2u * kEwPanel, kEwPanel);
            EwRing<DA> a;
            for (int i = 0; i < n_my; i++) {
                const uint32_t x = (uint32_t)(i & 1);
                tc::mbar_wait(&x0_full[x], (uint32_t)(i >> 1) & 1u);
                if (i >= DA) {
                    if (Cf::kWide) tc::mbar_wait(&d_free[a.slot], a.ph ^ 1u);  // the slot's previous tenant: D^T(i - 2), drained by EF
                    else tc::mbar_wait(&acc_free[a.slot], a.ph ^ 1u);
                }
                tc::fence_after_sync();
                const uint64_t ah = adv(xh0, x * 4u * kEwPanel), al = adv(xl0, x * 4u * kEwPanel);
                const uint32_t dacc = tmem + Cf::kAccCol + (uint32_t)a.slot * Cf::kAccStride;
                tc::mma_tf32(dacc, al, w0h, idesc, 0);
                tc::mma_tf32(dacc, ah, w0l, idesc, 1);
                tc::mma_tf32(dacc, ah, w0h, idesc, 1);
                tc::mma_commit(&s0_done[a.slot]);
                tc::mma_commit(&x0_free[x]);
                a.next();
            }
        } else if (warp == 17) {
            // =========================== MA: attention stage 1, transposed, weights from TMEM ===========================
            const uint32_t idesc = tc::make_idesc_tf32(128, 128);
            const uint64_t xh0 = tc::make_sdesc(sb + Cf::img, kEwPanel);
            const uint64_t xl0 = tc::make_sdesc(sb + Cf::img + (A0P / 4) * kEwPanel, kEwPanel);
            const uint32_t w_hi = tmem + Cf::kWCol, w_lo = w_hi + A0P;
            EwRing<DI> d;
            EwRing<DD> dd;
            for (int i = 0; i < n_my; i++) {
                if (i >= DD) tc::mbar_wait(&d_free[dd.slot], dd.ph ^ 1u);
                uint64_t bh = adv(xh0, (uint32_t)d.slot * Cf::kImgBytes), bl = adv(xl0, (uint32_t)d.slot * Cf::kImgBytes);
                const uint32_t dc = tmem + Cf::kDCol + (uint32_t)dd.slot * 128u;
#pragma unroll 1
                for (int h = 0; h < KH; h++) {
                    // (kWide: D^T shares the slot of S0 -- every K half must have left it before the first MMA)
                    for (int hw = (Cf::kWide ? 0 : h); hw < (Cf::kWide ? (h == 0 ? KH : 0) : h + 1); hw++)
                        tc::mbar_wait(&e0_done[d.slot * KH + hw], d.ph);
                    tc::fence_after_sync();
#pragma unroll 1
                    for (int ks = h * (A0P / 8 / KH); ks < (h + 1) * (A0P / 8 / KH); ks++) {
                        ew_mma_ts(dc, w_lo + (uint32_t)ks * 8u, bh, idesc, ks > 0);
                        ew_mma_ts(dc, w_hi + (uint32_t)ks * 8u, bl, idesc, 1);
                        ew_mma_ts(dc, w_hi + (uint32_t)ks * 8u, bh, idesc, 1);
                        bh = adv(bh, 2u * kEwPanel); bl = adv(bl, 2u * kEwPanel);
                    }
                    tc::mma_commit(&img_free[d.slot * KH + h]);
                }
                tc::mma_commit(&d_full[dd.slot]);
                d.next();
                dd.next();
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static bool edge_ws_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("GRIDGCN_EDGE_WS");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

template <int A0P>
static int launch_edge_ws_t(const TcParams &p, cudaStream_t st, int sms) {
    using Cf = EwCfg<A0P>;
    const ConvParams &c = p.c;
    const int nchunk = (c.Cout + 127) / 128;
    const int cpt = c.K >= 128 ? 1 : 128 / c.K;
    int log2k = 0;
    while ((1 << log2k) < c.K) log2k++;
    const long long centers = (long long)c.B * c.O;
    const long long units = (centers + cpt - 1) / cpt;
    if (units > 0x7fffffff) return -1;
    int dev = 0;
    cudaGetDevice(&dev);
    static PerDeviceOnce attr_set;
    const size_t smem = Cf::total + 1024;
    if (!attr_set.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(edge_ws_kernel<A0P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set.set(dev);
    }
    const long long owners = std::max<long long>(1, std::min<long long>(units, sms / nchunk));
    static const int blocked = [] { const char *e = getenv("GRIDGCN_EDGE_WS_BLOCKED"); return e ? atoi(e) : 1; }();
    edge_ws_kernel<A0P><<<(int)(owners * nchunk), kEwThreads, smem, st>>>(p, (int)units, cpt, log2k, nchunk, blocked);
    return (int)cudaGetLastError();
}

// Returns -1 when the layer does not fit this kernel (the caller falls back to edge_tc_kernel / edge_wide_kernel).
int launch_edge_ws(const TcParams &p, cudaStream_t st) {
    const ConvParams &c = p.c;
    if (!edge_ws_enabled()) return -1;
    if (p.nsplit != 3 || p.has_ff || !p.has_att || p.dbg != nullptr || c.Cin <= 0) return -1;
    if (!(c.K == 16 || c.K == 32 || c.K == 64 || c.K == 128)) return -1;
    if ((c.Cout & 127) != 0 || c.Cout > 1024 || p.a1.Cin != p.a0_cout) return -1;
    if ((long long)c.B * c.O >= (1LL << 31) || (long long)c.B * c.Nprev * c.Cout >= (1LL << 32)) return -1;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < (c.Cout + 127) / 128) return -1;
    switch (p.a1.Kp) {
        case 32: return launch_edge_ws_t<32>(p, st, sms);
        case 64: return launch_edge_ws_t<64>(p, st, sms);
        case 128: return launch_edge_ws_t<128>(p, st, sms);
        default: return -1;
    }
}

}  // namespace gg
