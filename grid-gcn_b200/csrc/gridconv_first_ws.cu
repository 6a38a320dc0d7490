// GridConv first layer (no input features: the per-edge feature MLP starts from the geo vector) as ONE
// persistent, warp-specialised tcgen05 pipeline per SM -- GRIDGCN_PRECISION_TF32X3.
//
// Reference semantics: segmentation/models/gcn_module_g_att.py:120-170 (verts_pair_func: feature MLP,
// attention MLP, product), :45-79 (max pool over the K slots), :172-287 (sub_g_update), gathers of
// utils/ops.py:78-93.  Same math as edge_first64_kernel (gridconv_tc.cu); what changes is the schedule:
//
//   * EVERY stage runs on the tensor core, also the tiny-K ones.  Stage 0 takes the 8-vector
//     [dist, dx, dy, dz, cx, cy, cz, 1] of an edge (K = 8: one k step) against a merged weight image whose rows
//     are the feature stage 0 (3 -> H0, bias in the "1" column) and the folded attention stage 0 (10 -> A0,
//     att0_folded_weight): D0[edge, H0p + A0p].  The CUDA cores only build that vector and run the
//     TMEM -> relu -> hi/lo split epilogues (r01 profile: the CUDA-core stages were a third of the 9500
//     warp-instructions per 128-edge tile and instruction issue, not the tensor pipe, was the limit).
//   * A operands from TENSOR MEMORY where it pays (measured, tools/issue_probe.cu + tools/ts_probe.cu: a
//     shared-memory-operand MMA is bound by ~124 B/cycle of operand fetch -- M=128 x N=32 takes 41 cycles instead
//     of 16 -- and one thread issues at most one MMA per 56 cycles):
//       - hidden feature stage: its A operand (the stage-0 activations, hi | lo) is written by the stage-0
//         epilogue straight into TMEM (tcgen05.st, thread = edge = lane), over the dead stage-0 accumulator;
//       - last feature stage (transposed, M = 64): its A operand, the weights, sits in TMEM for the whole kernel.
//   * Roles (25 warps, one CTA per SM, units of 128 edges round-robin over the CTAs):
//       E0   warps  4-7   stage-0 epilogue (thread = edge): D0 -> relu -> hi/lo: xf -> TMEM, xa -> smem
//       EH   warps  8-11  hidden-stage epilogue (thread = edge): relu(D + b) -> hi/lo image xf2 (smem)
//       G    warps 21-24  gather: neighbour index -> table row (128-bit) + centre -> input image X0 (smem)
//       EF   warps 12-15 and 0-3: final epilogue (thread = channel x 64-edge half, each group 32 of the 64 columns):
//                         relu(F + b) * relu(G + b), max over K, store -- the role that bounds the kernel
//       MS / MH / MA / MF0 / MF1   warps 16-20: one thread each issues the MMAs of stage 0 / the hidden stage /
//            attention stage 1 / the last feature stage (one warp per 64-edge half) and commits their mbarriers
//     Every hand-over is an mbarrier (ONE arrival per epilogue warp -- every arrival wakes the warps sleeping on
//     any barrier, r02f: 128 thread arrivals per hand-over made half of all executed instructions re-polls --,
//     tcgen05.commit from the MMA threads); every resource has its own barrier ring (a wait on a barrier of a ring of a DIFFERENT depth can
//     alias phases and deadlock).
//   * Last stages transposed with M = 64: D^T[ch, edge]; the 64 edges [64h, 64h+64) of a unit go to TMEM lanes
//     32q+16h .. +15 (cta_group::1 M = 64 layout, tools/m64_probe.cu; A and D must use the same lane half),
//     F in columns [0,64), G in [64,128) of the SAME lane: the product needs no shuffle and the max over a centre's
//     K consecutive edges is a run of FMNMX in one thread.
//
// TMEM map (512 columns): front slot a (2 x 96): [0,48) stage-0 accumulator, overwritten by xf hi [0,32) | lo
// [32,64); [64,96) hidden accumulator.  F|G slot f (2 x 128) at 192.  Last-stage weights hi | lo at 448.
// Every wait is bounded (tc::mbar_wait traps) so a protocol error fails the launch instead of hanging.
#include "gridconv_tc.cuh"

#include <cstdlib>

// -DGG_WS_TIMING (tools/build_variant.py): CTA 0 accumulates, per role, the cycles spent in every barrier wait and the
// role's total loop time into the debug buffer (gridgcn_debug_phase_buffer, 32 x u64; tools/ws_timing.py).
#ifdef GG_WS_TIMING
#define WS_WAIT(idx, bar, par)                                        \
    {                                                                 \
        const long long t_ = clock64();                               \
        tc::mbar_wait(bar, par);                                      \
        if (ws_timing) ws_tw[idx] += (unsigned long long)(clock64() - t_); \
    }
#else
#define WS_WAIT(idx, bar, par) tc::mbar_wait(bar, par)
#endif

namespace gg {

constexpr int kWsThreads = 800;
constexpr int kDX = 2;  // X0 input images                      (smem, 8 KB each)
constexpr int kDA = 2;  // front slots: S0 acc / xf / hidden acc (TMEM, 96 columns each: [0, 192))
constexpr int kDI = 3;  // activation images xf2 | xa            (smem)
constexpr int kDF = 2;  // F | G accumulators                    (TMEM, 128 columns each: [192, 448))
constexpr uint32_t kFrontCols = 96, kXfLoCol = 32, kHidCol = 64, kFgCol0 = 192, kWffCol = 448;
constexpr uint32_t kPanel = 2048;  // one [128 rows x 4 k] panel of a K-major image
static_assert(kDX == kDA, "G waits for the X0 slot on the stage-0 barrier of the same ring position");

template <int D>
struct RingPos {  // slot and use-count parity of the current unit in a ring of depth D
    int slot = 0;
    uint32_t ph = 0;
    __device__ __forceinline__ void next() {
        if (++slot == D) { slot = 0; ph ^= 1u; }
    }
};

struct FirstWsLayout {
    int H0p, A0p, N0, H1n, H1p;
    uint32_t w0_hi, w0_lo, wfh, wa1_hi, wa1_lo, bias_h, x0, img, img_stride, xf_lo, xa_hi, xa_lo, total;
};

__host__ __device__ inline FirstWsLayout first_ws_layout(const TcParams &p) {
    FirstWsLayout L;
    L.H0p = pad_to(p.f0_cout, 8);
    L.A0p = p.a1.Kp;
    L.N0 = pad_to(L.H0p + L.A0p, 16);
    L.H1n = p.fh[0].Np;
    L.H1p = p.ff.Kp;
    uint32_t o = 0;
    L.w0_hi = o; o += (uint32_t)L.N0 * 32u;
    L.w0_lo = o; o += (uint32_t)L.N0 * 32u;
    L.wfh = o;   o += 2u * (uint32_t)p.fh[0].Np * (uint32_t)p.fh[0].Kp * 4u;
    L.wa1_hi = o; o += (uint32_t)(L.A0p / 4) * 1024u;
    L.wa1_lo = o; o += (uint32_t)(L.A0p / 4) * 1024u;
    L.bias_h = o; o += 64u * 4u;
    o = (o + 127u) & ~127u;
    L.x0 = o; o += (uint32_t)kDX * 4u * kPanel;  // per slot: hi panels 0-1, lo panels 0-1
    L.img = o;                                   // per slot: xf2 hi | xf2 lo | xa hi | xa lo
    L.xf_lo = (uint32_t)(L.H1p / 4) * kPanel;
    L.xa_hi = 2u * L.xf_lo;
    L.xa_lo = L.xa_hi + (uint32_t)(L.A0p / 4) * kPanel;
    L.img_stride = L.xa_lo + (uint32_t)(L.A0p / 4) * kPanel;
    o += (uint32_t)kDI * L.img_stride;
    L.total = o;
    return L;
}

// D[tmem] (+)= A[tmem] * B[smem]^T: A operand from tensor memory (row = lane, k = column), issued by ONE thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 registers -> 16 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

// relu -> hi/lo split (hardware truncation model, tc_common.cuh split_op<3>) of four accumulator columns,
// one 16-byte store per image
__device__ __forceinline__ void relu_split_store4(const uint32_t *v, const float4 b, uint8_t *dst_hi, uint8_t *dst_lo) {
    float x[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])}, lo[4];
    tc::add2(x[0], x[1], b.x, b.y);  // packed pairs (FADD2): bit-identical to the scalar operations
    tc::add2(x[2], x[3], b.z, b.w);
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = fmaxf(x[i], 0.f);
    tc::split_lo2(x[0], x[1], lo[0], lo[1]);
    tc::split_lo2(x[2], x[3], lo[2], lo[3]);
    *reinterpret_cast<float4 *>(dst_hi) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4 *>(dst_lo) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}
__device__ __forceinline__ void relu_split_store4_nobias(const uint32_t *v, uint8_t *dst_hi, uint8_t *dst_lo) {
    float x[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = fmaxf(__uint_as_float(v[i]), 0.f);
    tc::split_lo2(x[0], x[1], lo[0], lo[1]);
    tc::split_lo2(x[2], x[3], lo[2], lo[3]);
    *reinterpret_cast<float4 *>(dst_hi) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4 *>(dst_lo) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

template <int H0P, int A0P, int H1P>
__global__ void __launch_bounds__(kWsThreads, 1)
edge_first_ws_kernel(const __grid_constant__ TcParams p, int num_units, int cpt, int log2k) {
    constexpr int N0 = (H0P + A0P + 15) / 16 * 16;
    static_assert(H0P % 16 == 0 && N0 <= 64 && 2 * H0P <= 64 && 2 * H1P <= 64 && H0P % 8 == 0 && H1P % 8 == 0 && A0P % 8 == 0, "TMEM map");
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[kDX + 4 * kDA + 2 * kDI + 3 * kDF];
    __shared__ uint32_t tmem_base_s;
    __shared__ float ef_pair[4][32];  // K >= 64: partial maxima handed from the second final-epilogue group to the first
    uint64_t *x0_full = bars;                                                    // G -> MS (4: one arrival per warp)
    uint64_t *s0_done = x0_full + kDX, *e0_done = s0_done + kDA;                 // MS -> E0, G (1); E0 -> MH (4: one arrival per warp)
    uint64_t *h_done = e0_done + kDA, *acc_free = h_done + kDA;                  // MH -> EH (1); EH -> MS (4: one arrival per warp)
    uint64_t *eh_done = acc_free + kDA, *img_free = eh_done + kDI;               // EH -> MA, MF0, MF1 (4: one arrival per warp); MA + MF0 + MF1 -> E0 (3)
    uint64_t *f_full = img_free + kDI, *g_full = f_full + kDF, *fg_free = g_full + kDF;  // MF0 + MF1 -> EF (2); MA -> EF (1); EF -> MA, MF (4: one arrival per warp)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ConvParams &c = p.c;
    const int C = c.Cout;
    const FirstWsLayout L = first_ws_layout(p);

    // ---- one-time set-up: TMEM, barriers, resident weight images ----
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 32) {
        for (int i = 0; i < kDX; i++) tc::mbar_init(&x0_full[i], 4);
        for (int i = 0; i < kDA; i++) {
            tc::mbar_init(&s0_done[i], 1);
            tc::mbar_init(&e0_done[i], 4);
            tc::mbar_init(&h_done[i], 1);
            tc::mbar_init(&acc_free[i], 4);
        }
        for (int i = 0; i < kDI; i++) {
            tc::mbar_init(&eh_done[i], 4);
            tc::mbar_init(&img_free[i], 3);
        }
        for (int i = 0; i < kDF; i++) {
            tc::mbar_init(&f_full[i], 2);
            tc::mbar_init(&g_full[i], 1);
            tc::mbar_init(&fg_free[i], 8);
        }
        tc::mbar_init_fence();
    }
    {
        // merged stage-0 image W0[N0 x 8]: rows [0, H0p) feature stage 0 = (0, w_x, w_y, w_z, 0, 0, 0, b),
        // rows [H0p, H0p + A0p) folded attention stage 0 = (w_dist, W_d + W_n | W_c + W_n, b)
        const uint32_t lbo0 = (uint32_t)N0 * 16u;
        for (int i = tid; i < N0 * 8; i += kWsThreads) {
            const int n = i >> 3, k = i & 7;
            float w = 0.f;
            if (n < H0P) {
                if (n < p.f0_cout) {
                    if (k >= 1 && k <= 3) w = __ldg(p.f0_w + (size_t)n * 3 + (k - 1));
                    else if (k == 7) w = __ldg(p.f0_b + n);
                }
            } else if (n < H0P + A0P) {
                w = att0_folded_weight(p.a0_w, p.a0_b, p.a0_cin, p.a0_cout, n - H0P, k);
            }
            float hi, lo;
            tc::split_tf32(w, hi, lo);
            const uint32_t off = tc::kmajor_off((uint32_t)n, (uint32_t)k, lbo0);
            *reinterpret_cast<float *>(smem + L.w0_hi + off) = hi;
            *reinterpret_cast<float *>(smem + L.w0_lo + off) = lo;
        }
        {   // hidden stage: packed plain image (hi then lo), as is (B operand)
            const TcStage &st = p.fh[0];
            const int n4 = 2 * st.Np * st.Kp / 4;
            const float4 *src = reinterpret_cast<const float4 *>(p.packed + st.w_off);
            float4 *dst = reinterpret_cast<float4 *>(smem + L.wfh);
            for (int i = tid; i < n4; i += kWsThreads) dst[i] = __ldg(src + i);
        }
        {   // attention stage 1 (A operand from shared memory): rows 0..63 of every [128 x 4] panel of the packed image
            const TcStage &st = p.a1;
            const int per_img = (st.Kp / 4) * 256;
            for (int i = tid; i < 2 * per_img; i += kWsThreads) {
                const int img = i / per_img, r = i % per_img, P = r >> 8, w = r & 255;
                const int t = P / (kSliceK / 4), pp = P % (kSliceK / 4);
                const int kw = min(kSliceK, st.Kp - t * kSliceK);
                const float *src = p.packed + st.w_off + (size_t)t * 2 * 128 * kSliceK + (img ? 128 * kw : 0) + pp * 512 + w;
                reinterpret_cast<float *>(smem + (img ? L.wa1_lo : L.wa1_hi))[r] = __ldg(src);
            }
        }
        for (int i = tid; i < 64; i += kWsThreads)
            reinterpret_cast<float *>(smem + L.bias_h)[i] = i < p.fh[0].Cout ? __ldg(p.fh[0].bias + i) : 0.f;
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t q = (uint32_t)(warp & 3);
    const uint32_t lane_base = (q * 32u) << 16;                             // TMEM lane quadrant of this warp
    if (warp < 4) {
        // last feature stage weights -> TMEM (A operand, M = 64 layout: row 16q + i in lanes 32q + i and 32q + 16 + i,
        // one copy per 64-edge half), hi in columns [kWffCol, +H1P), lo in [kWffCol + 32, +H1P)
        const int m = 16 * (int)q + (lane & 15);
        const float *wrow = c.w[c.n_feat - 1] + (size_t)m * p.ff.Cin;
#pragma unroll
        for (int k0 = 0; k0 < H1P; k0 += 8) {
            float hi[8], lo[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float w = (m < p.ff.Cout && k0 + k < p.ff.Cin) ? __ldg(wrow + k0 + k) : 0.f;
                tc::split_tf32(w, hi[k], lo[k]);
            }
            tmem_st8(tmem + lane_base + kWffCol + (uint32_t)k0, hi);
            tmem_st8(tmem + lane_base + kWffCol + 32u + (uint32_t)k0, lo);
        }
        tc::tmem_st_wait();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    const int n_my = (int)blockIdx.x < num_units ? (num_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const unsigned centers_total = (unsigned)c.B * (unsigned)c.O;
    const int out_w = 4 + C;
#ifdef GG_WS_TIMING
    const bool ws_timing = p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && ((warp < 16 && (warp & 3) == 0) || (warp >= 16 && warp <= 21));
    unsigned long long ws_tw[16];
    for (int i = 0; i < 16; i++) ws_tw[i] = 0;
    const long long ws_t0 = clock64();
#endif
    const uint32_t row = q * 32u + (uint32_t)lane;                          // edge row of the unit (E0, EH)
    const uint32_t row_off = (row >> 3) * 128u + (row & 7u) * 16u;           // its offset inside a panel

    if (warp >= 21) {
        // =========================== G: gather + input image ===========================
        // (its own four warps: the fence.proxy.async behind every image store waits for the prefetch loads of the
        //  following units, ~1500 cycles -- harmless here, fatal on the critical path of an epilogue role, r02)
        const int r = tid - 672;
        const uint32_t row_off = ((uint32_t)r >> 3) * 128u + ((uint32_t)r & 7u) * 16u;
        const int my_cl = r >> log2k, my_slot = r & ((1 << log2k) - 1);
        const int Nprev = c.Nprev, O = c.O;
        const int rows_total = c.B * c.Nprev;
        const int row_w = 4 + c.Cin;
        auto center_of = [&](int i) -> unsigned { return (unsigned)(blockIdx.x + i * gridDim.x) * (unsigned)cpt + (unsigned)my_cl; };
        auto load_idx = [&](int i, bool &valid) -> int {
            const unsigned center = center_of(i);
            valid = i < n_my && my_cl < cpt && center < centers_total;
            return valid ? __ldg(c.nebidx + ((size_t)center << log2k) + my_slot) : 0;
        };
        auto load_row = [&](int i, bool valid, int idx, float4 &head, float4 &cent) {
            head = make_float4(0.f, 0.f, 0.f, 0.f);
            cent = head;
            if (valid) {
                const unsigned center = center_of(i);
                const int b = (int)(center / (unsigned)O);
                int gi = idx + b * Nprev;  // take(): clip after the batch offset (utils/ops.py:90-92)
                gi = gi < 0 ? 0 : (gi >= rows_total ? rows_total - 1 : gi);
                const float *src = c.table + (size_t)gi * row_w;
                if ((row_w & 3) == 0) head = __ldg(reinterpret_cast<const float4 *>(src));
                else head = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 0.f);
                cent = __ldg(c.cent + center);
            }
        };
        bool v_cur, v_nxt;
        int idx_nxt;
        float4 head, cent;
        {
            const int idx0 = load_idx(0, v_cur);
            load_row(0, v_cur, idx0, head, cent);
            idx_nxt = load_idx(1, v_nxt);
        }
        RingPos<kDA> sa;  // front-slot position of unit i - kDX (whose stage-0 MMAs read this X0 slot)
        for (int i = 0; i < n_my; i++) {
            const int x = i & 1;
            // requests for the following units first: row / centre of unit i+1, index of unit i+2
            float4 head_n, cent_n;
            load_row(i + 1, v_nxt, idx_nxt, head_n, cent_n);
            bool v_n2;
            const int idx_n2 = load_idx(i + 2, v_n2);
            // input vector of this edge
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
            if (i >= kDX) {
                WS_WAIT(0, &s0_done[sa.slot], sa.ph);
                sa.next();
            }
            uint8_t *x0 = smem + L.x0 + (uint32_t)x * 4u * kPanel + row_off;
            *reinterpret_cast<float4 *>(x0) = make_float4(in[0], in[1], in[2], in[3]);
            *reinterpret_cast<float4 *>(x0 + kPanel) = make_float4(in[4], in[5], in[6], in[7]);
            *reinterpret_cast<float4 *>(x0 + 2 * kPanel) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4 *>(x0 + 3 * kPanel) = make_float4(lo[4], lo[5], lo[6], lo[7]);
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&x0_full[x]);
            // centre columns of the output rows
            if (r < cpt * 4) {
                const unsigned center = (unsigned)(blockIdx.x + i * gridDim.x) * (unsigned)cpt + (unsigned)(r >> 2);
                if (center < centers_total)
                    c.out[(size_t)center * out_w + (r & 3)] = __ldg(reinterpret_cast<const float *>(c.cent + center) + (r & 3));
            }
            head = head_n; cent = cent_n; v_cur = v_nxt;
            idx_nxt = idx_n2; v_nxt = v_n2;
        }
    } else if (warp >= 8 && warp < 12) {
        // =========================== EH: hidden-stage epilogue ===========================
        const float *bias_h = reinterpret_cast<const float *>(smem + L.bias_h);
        RingPos<kDA> a;
        RingPos<kDI> d;
        for (int i = 0; i < n_my; i++) {
            WS_WAIT(3, &h_done[a.slot], a.ph);
            tc::fence_after_sync();
            const uint32_t taddr = tmem + lane_base + (uint32_t)a.slot * kFrontCols + kHidCol;
            uint8_t *img = smem + L.img + (uint32_t)d.slot * L.img_stride + row_off;
            constexpr int H1L = (H1P + 15) / 16 * 16;
            uint32_t v[H1L];
#pragma unroll
            for (int c0 = 0; c0 < H1L; c0 += 16) tc::tmem_ld16(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(v + c0));
            tc::tmem_ld_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_free[a.slot]);  // the whole front slot (xf and both accumulators) is dead now
#pragma unroll
            for (int cc = 0; cc < H1P; cc += 4) {
                uint8_t *dst = img + (uint32_t)(cc >> 2) * kPanel;
                relu_split_store4(v + cc, *reinterpret_cast<const float4 *>(bias_h + cc), dst, dst + L.xf_lo);
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&eh_done[d.slot]);
            a.next();
            d.next();
        }
    } else if (warp >= 4 && warp < 8) {
        // =========================== E0: stage-0 epilogue ===========================
        // D0 -> relu -> hi/lo.  The feature half goes back into TMEM over the accumulator (A operand of the hidden
        // stage: row = this thread's lane), the attention half into the shared-memory image (B operand of stage a1).
        RingPos<kDA> a;
        RingPos<kDI> d;
        for (int i = 0; i < n_my; i++) {
            if (i >= kDI) WS_WAIT(1, &img_free[d.slot], d.ph ^ 1u);  // last-stage MMAs of unit i-3 have read the images
            WS_WAIT(2, &s0_done[a.slot], a.ph);
            tc::fence_after_sync();
            const uint32_t taddr = tmem + lane_base + (uint32_t)a.slot * kFrontCols;
            uint8_t *img = smem + L.img + (uint32_t)d.slot * L.img_stride + row_off;
            {   // attention half first: its columns are overwritten by the lo part of xf below
                uint32_t va[16];
                if (A0P > 8) tc::tmem_ld16(taddr + (uint32_t)H0P, va);
                else tmem_ld8(taddr + (uint32_t)H0P, va);
                tc::tmem_ld_wait();
#pragma unroll
                for (int cc = 0; cc < A0P; cc += 4) {
                    uint8_t *dst = img + L.xa_hi + (uint32_t)(cc >> 2) * kPanel;
                    relu_split_store4_nobias(va + cc, dst, dst + (L.xa_lo - L.xa_hi));
                }
            }
#pragma unroll
            for (int c0 = 0; c0 < H0P; c0 += 16) {
                uint32_t v[16];
                tc::tmem_ld16(taddr + (uint32_t)c0, v);
                tc::tmem_ld_wait();
                float xh[16], xl[16];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    xh[j] = fmaxf(__uint_as_float(v[j]), 0.f);
                    xh[j + 1] = fmaxf(__uint_as_float(v[j + 1]), 0.f);
                    tc::split_lo2(xh[j], xh[j + 1], xl[j], xl[j + 1]);
                }
                tmem_st16(taddr + (uint32_t)c0, xh);
                tmem_st16(taddr + kXfLoCol + (uint32_t)c0, xl);
            }
            tc::tmem_st_wait();
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&e0_done[a.slot]);
            a.next();
            d.next();
        }
    } else if (warp < 4 || (warp >= 12 && warp < 16)) {
        // =========================== EF: final epilogue (two groups of four warps) ===========================
        // lane l of a warp with quadrant q: channel 16q + (l & 15), edges [64h, 64h + 64) of the unit with h = l >> 4;
        // group g (warps 12-15: 0, warps 0-3: 1) reduces the 32 columns [32g, 32g + 32) of those.
        // (Tried, r02: starting the accumulators at the bias through tcgen05.st.x16 -- the 16 source registers of
        // every store have to be filled with MOVs, which costs what the bias adds cost.)
        const int grp = warp < 4 ? 1 : 0;
        const int h = lane >> 4;
        const int ch = 16 * (int)q + (lane & 15);
        const bool chv = ch < C;
        const float bf = chv ? __ldg(p.ff.bias + ch) : 0.f;
        const float bg = chv ? __ldg(p.a1.bias + ch) : 0.f;
        float *out_ch = c.out + 4 + ch;
        const int kmask = (1 << log2k) - 1;
        RingPos<kDF> f;
        for (int i = 0; i < n_my; i++) {
            const unsigned c_base = (unsigned)(blockIdx.x + i * gridDim.x) * (unsigned)cpt;
            WS_WAIT(4, &f_full[f.slot], f.ph);
            WS_WAIT(5, &g_full[f.slot], f.ph);
            tc::fence_after_sync();
            const uint32_t taddr = tmem + lane_base + kFgCol0 + (uint32_t)f.slot * 128u + (uint32_t)(32 * grp);
            float m = -3.402823466e+38f;
            uint32_t fa[16], ga[16];
            auto reduce16 = [&](const uint32_t (&fv)[16], const uint32_t (&gv)[16], int c0) {
                float pr[16];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    // :167 att * feats.  relu(F) >= 0, so relu(F) * relu(G) == max(relu(F) * G, 0): the attention ReLU is
                    // hoisted out of the max over K (applied once per centre below).  Packed fp32 pairs (FADD2 / FMUL2).
                    float x0 = __uint_as_float(fv[j]), x1 = __uint_as_float(fv[j + 1]);
                    float a0 = __uint_as_float(gv[j]), a1 = __uint_as_float(gv[j + 1]);
                    tc::add2(x0, x1, bf, bf);
                    tc::add2(a0, a1, bg, bg);
                    x0 = fmaxf(x0, 0.f);
                    x1 = fmaxf(x1, 0.f);
                    tc::mul2(x0, x1, a0, a1);
                    pr[j] = x0;
                    pr[j + 1] = x1;
                }
                float mm = pr[0];
#pragma unroll
                for (int j = 1; j < 16; j++) mm = fmaxf(mm, pr[j]);
                m = fmaxf(m, mm);
                const int e_end = 64 * h + 32 * grp + c0 + 16;  // edges of this lane reduced so far end here
                if (log2k <= 5 && (e_end & kmask) == 0) {       // K <= 32: the centre lies inside this group's columns
                    const unsigned center = c_base + (unsigned)((e_end >> log2k) - 1);
                    if (chv && center < centers_total)
                        out_ch[(size_t)center * out_w] = fmaxf(m, 0.f) * __ldg(c.centmsk + center);
                    m = -3.402823466e+38f;
                }
            };
            tc::tmem_ld16(taddr, fa);
            tc::tmem_ld16(taddr + 64u, ga);
            tc::tmem_ld_wait();
            reduce16(fa, ga, 0);
            tc::tmem_ld16(taddr + 16u, fa);
            tc::tmem_ld16(taddr + 80u, ga);
            tc::tmem_ld_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&fg_free[f.slot]);  // every TMEM read of this warp is done (8 arrivals free the slot)
            reduce16(fa, ga, 16);
            if (log2k >= 6) {  // K = 64 / 128: the centre spans both groups' columns -- group 1 hands its partial maximum over
                if (grp == 1) ef_pair[q][lane] = m;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)q) : "memory");
                if (grp == 0) {
                    m = fmaxf(m, ef_pair[q][lane]);
                    if (log2k == 7) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));  // one centre per unit: both lane halves
                    const unsigned center = c_base + (log2k == 7 ? 0u : (unsigned)h);
                    if (chv && center < centers_total && (log2k == 6 || h == 0))
                        out_ch[(size_t)center * out_w] = fmaxf(m, 0.f) * __ldg(c.centmsk + center);
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + (int)q) : "memory");
            }
            f.next();
        }
    } else if (lane == 0) {
        // =========================== MMA issue: one thread per stage ===========================
        // Descriptors are built once; per use only the 16-byte-unit start-address field (bits 0-13, never
        // overflowing: shared addresses < 256 KB) is advanced by an addition.
        const uint32_t sb = tc::smem_u32(smem);
        auto adv = [](uint64_t dsc, uint32_t bytes) -> uint64_t { return dsc + (uint64_t)(bytes >> 4); };
        if (warp == 16) {
            // ---- MS: stage 0, D0[edge, H0p + A0p], K = 8, both operands from shared memory.  The front slot held
            //      unit i - kDA: wait for its hidden epilogue ----
            const uint32_t idesc = tc::make_idesc_tf32(128, N0);
            const uint32_t lbo_w0 = (uint32_t)N0 * 16u;
            const uint64_t w0h = tc::make_sdesc(sb + L.w0_hi, lbo_w0), w0l = tc::make_sdesc(sb + L.w0_lo, lbo_w0);
            const uint64_t xh0 = tc::make_sdesc(sb + L.x0, kPanel), xl0 = tc::make_sdesc(sb + L.x0 + 2u * kPanel, kPanel);
            RingPos<kDX> x;
            RingPos<kDA> a;
            for (int i = 0; i < n_my; i++) {
                WS_WAIT(6, &x0_full[x.slot], x.ph);
                if (i >= kDA) WS_WAIT(7, &acc_free[a.slot], a.ph ^ 1u);
                tc::fence_after_sync();
                const uint64_t ah = adv(xh0, (uint32_t)x.slot * 4u * kPanel), al = adv(xl0, (uint32_t)x.slot * 4u * kPanel);
                const uint32_t dacc = tmem + (uint32_t)a.slot * kFrontCols;
                tc::mma_tf32(dacc, al, w0h, idesc, 0);
                tc::mma_tf32(dacc, ah, w0l, idesc, 1);
                tc::mma_tf32(dacc, ah, w0h, idesc, 1);
                tc::mma_commit(&s0_done[a.slot]);
                x.next();
                a.next();
            }
        } else if (warp == 17) {
            // ---- MH: hidden feature stage D[edge, ch], A = xf hi | lo in TMEM (row = lane), B = weights in smem ----
            const uint32_t idesc = tc::make_idesc_tf32(128, L.H1n);
            const uint32_t lbo_wh = (uint32_t)L.H1n * 16u;
            const uint64_t wh0 = tc::make_sdesc(sb + L.wfh, lbo_wh);
            const uint64_t wl0 = tc::make_sdesc(sb + L.wfh + (uint32_t)L.H1n * (uint32_t)H0P * 4u, lbo_wh);
            constexpr int ks_h = H0P / 8;
            RingPos<kDA> a;
            for (int i = 0; i < n_my; i++) {
                WS_WAIT(8, &e0_done[a.slot], a.ph);
                tc::fence_after_sync();
                const uint32_t a_hi = tmem + (uint32_t)a.slot * kFrontCols, a_lo = a_hi + kXfLoCol, dacc = a_hi + kHidCol;
                uint64_t bh = wh0, bl = wl0;
#pragma unroll
                for (int ks = 0; ks < ks_h; ks++) {
                    mma_tf32_ts(dacc, a_lo + (uint32_t)ks * 8u, bh, idesc, ks > 0);
                    mma_tf32_ts(dacc, a_hi + (uint32_t)ks * 8u, bl, idesc, 1);
                    mma_tf32_ts(dacc, a_hi + (uint32_t)ks * 8u, bh, idesc, 1);
                    bh = adv(bh, 2u * lbo_wh); bl = adv(bl, 2u * lbo_wh);
                }
                tc::mma_commit(&h_done[a.slot]);
                a.next();
            }
        } else if (warp == 18) {
            // ---- MA: attention stage 1 -> G, transposed, M = 64 channels x N = 64 edges per half, both operands from
            //      shared memory; the two halves' MMAs alternate.  Starts with the last feature stage (the F|G slot is
            //      then held for the shortest time). ----
            const uint32_t idesc = tc::make_idesc_tf32(64, 64);
            const uint64_t wh0 = tc::make_sdesc(sb + L.wa1_hi, 1024u), wl0 = tc::make_sdesc(sb + L.wa1_lo, 1024u);
            const uint64_t xh0 = tc::make_sdesc(sb + L.img + L.xa_hi, kPanel), xl0 = tc::make_sdesc(sb + L.img + L.xa_lo, kPanel);
            constexpr int nks = A0P / 8;
            RingPos<kDI> d;
            RingPos<kDF> f;
            for (int i = 0; i < n_my; i++) {
                WS_WAIT(9, &eh_done[d.slot], d.ph);
                if (i >= kDF) WS_WAIT(10, &fg_free[f.slot], f.ph ^ 1u);
                tc::fence_after_sync();
                uint64_t ah = wh0, al = wl0;
                uint64_t bh = adv(xh0, (uint32_t)d.slot * L.img_stride), bl = adv(xl0, (uint32_t)d.slot * L.img_stride);
                const uint32_t dc0 = tmem + kFgCol0 + 64u + (uint32_t)f.slot * 128u, dc1 = dc0 + (16u << 16);
#pragma unroll
                for (int ks = 0; ks < nks; ks++) {
                    const uint64_t bh1 = adv(bh, 1024u), bl1 = adv(bl, 1024u);  // rows 64..127 of the image
                    tc::mma_tf32(dc0, al, bh, idesc, ks > 0);
                    tc::mma_tf32(dc1, al, bh1, idesc, ks > 0);
                    tc::mma_tf32(dc0, ah, bl, idesc, 1);
                    tc::mma_tf32(dc1, ah, bl1, idesc, 1);
                    tc::mma_tf32(dc0, ah, bh, idesc, 1);
                    tc::mma_tf32(dc1, ah, bh1, idesc, 1);
                    ah = adv(ah, 2048u); al = adv(al, 2048u);
                    bh = adv(bh, 2u * kPanel); bl = adv(bl, 2u * kPanel);
                }
                tc::mma_commit(&g_full[f.slot]);
                tc::mma_commit(&img_free[d.slot]);
                d.next();
                f.next();
            }
        } else {
            // ---- MF0 / MF1 (warps 19, 20): last feature stage -> F for the 64-edge half h.  Transposed, M = 64:
            //      A = the weights in TMEM (lanes 32q + 16h + i), B = rows [64h, 64h + 64) of the xf2 image ----
            const uint32_t hh = (uint32_t)(warp - 19);
            const uint32_t idesc = tc::make_idesc_tf32(64, 64);
            const uint64_t xh0 = tc::make_sdesc(sb + L.img + hh * 1024u, kPanel), xl0 = tc::make_sdesc(sb + L.img + L.xf_lo + hh * 1024u, kPanel);
            const uint32_t w_hi = tmem + ((hh * 16u) << 16) + kWffCol, w_lo = w_hi + 32u;
            constexpr int nks = H1P / 8;
            RingPos<kDI> d;
            RingPos<kDF> f;
            for (int i = 0; i < n_my; i++) {
                WS_WAIT(11, &eh_done[d.slot], d.ph);
                if (i >= kDF) WS_WAIT(12, &fg_free[f.slot], f.ph ^ 1u);
                tc::fence_after_sync();
                uint64_t bh = adv(xh0, (uint32_t)d.slot * L.img_stride), bl = adv(xl0, (uint32_t)d.slot * L.img_stride);
                const uint32_t dc = tmem + ((hh * 16u) << 16) + kFgCol0 + (uint32_t)f.slot * 128u;
#pragma unroll
                for (int ks = 0; ks < nks; ks++) {
                    mma_tf32_ts(dc, w_lo + (uint32_t)ks * 8u, bh, idesc, ks > 0);
                    mma_tf32_ts(dc, w_hi + (uint32_t)ks * 8u, bl, idesc, 1);
                    mma_tf32_ts(dc, w_hi + (uint32_t)ks * 8u, bh, idesc, 1);
                    bh = adv(bh, 2u * kPanel); bl = adv(bl, 2u * kPanel);
                }
                tc::mma_commit(&f_full[f.slot]);
                tc::mma_commit(&img_free[d.slot]);
                d.next();
                f.next();
            }
        }
    }
#ifdef GG_WS_TIMING
    if (ws_timing) {
        const int role = warp < 16 ? warp >> 2 : (warp < 21 ? 4 + (warp - 16) : 9);  // EF1 E0 EH EF0 MS MH MA MF0 MF1 G
        for (int i = 0; i < 16; i++)
            if (ws_tw[i]) p.dbg[i] = ws_tw[i];
        p.dbg[16 + role] = (unsigned long long)(clock64() - ws_t0);
        if (role == 0) p.dbg[31] = (unsigned long long)n_my;
    }
#endif
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// Host side ------------------------------------------------------------------------------------------
static bool first_ws_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("GRIDGCN_FIRST_WS");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

// Returns -1 when the layer does not fit this kernel (the caller falls back to edge_first64_kernel /
// edge_tc_kernel), else the CUDA error code of the launch.
int launch_first_ws(const TcParams &p, cudaStream_t st) {
    const ConvParams &c = p.c;
    if (!first_ws_enabled()) return -1;
    #ifndef GG_WS_TIMING
    if (p.dbg != nullptr) return -1;
#endif
    if (p.nsplit != 3 || !p.has_ff || !p.has_att || !p.f0_cuda || p.nfh != 1) return -1;
    if (!(c.K == 16 || c.K == 32 || c.K == 64 || c.K == 128)) return -1;
    if (c.Cout > 64 || p.ff.Np != 128 || p.a1.Np != 128) return -1;
    const FirstWsLayout L = first_ws_layout(p);
    if (L.N0 > 64 || L.H1n > 32 || L.H1p > 32 || L.H0p > 32 || p.fh[0].Kp != L.H0p || p.fh[0].Cout > 32) return -1;
    if ((long long)c.B * c.O >= (1LL << 31) || (long long)c.B * c.Nprev >= (1LL << 31)) return -1;
    const size_t smem = (size_t)L.total + 1024;
    if (smem > 224 * 1024) return -1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // instantiated width combinations (H0p, A0p, H1p); anything else takes the older kernels
    void (*kern)(TcParams, int, int, int) = nullptr;
    int which = -1;
    if (L.H0p == 32 && L.A0p == 16 && L.H1p == 32) { kern = edge_first_ws_kernel<32, 16, 32>; which = 0; }
    else if (L.H0p == 16 && L.A0p == 16 && L.H1p == 32) { kern = edge_first_ws_kernel<16, 16, 32>; which = 1; }
    else if (L.H0p == 32 && L.A0p == 8 && L.H1p == 32) { kern = edge_first_ws_kernel<32, 8, 32>; which = 2; }
    if (!kern || L.H1n != pad_to(L.H1p, 16)) return -1;
    static PerDeviceOnce attr_set[3];
    if (!attr_set[which].done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set[which].set(dev);
    }
    const int cpt = c.K >= 128 ? 1 : 128 / c.K;
    int log2k = 0;
    while ((1 << log2k) < c.K) log2k++;
    const long long centers = (long long)c.B * c.O;
    const long long units = (centers + cpt - 1) / cpt;
    if (units > 0x7fffffff) return -1;
    const int blocks = (int)std::min<long long>(units, sms);
    kern<<<blocks, kWsThreads, smem, st>>>(p, (int)units, cpt, log2k);
    return (int)cudaGetLastError();
}

}  // namespace gg
