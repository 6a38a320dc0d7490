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
//     TMEM -> relu -> hi/lo split -> shared-memory epilogues (r01 profile: the CUDA-core stages were a third
//     of the 9500 warp-instructions per 128-edge tile and instruction issue, not the tensor pipe, was the limit).
//   * Roles (20 warps, one CTA per SM, units of 128 edges round-robin over the CTAs):
//       G   warps  0-3   gather: neighbour index -> table row (128-bit) + centre -> input image X0
//       E0  warps  4-7   epilogue of stage 0   (thread = edge): D0 -> relu -> hi/lo images xf (H0p) | xa (A0p)
//       EH  warps  8-11  epilogue of the hidden feature stage (thread = edge): relu(D + b) -> xf in place
//       EF  warps 12-15  final epilogue (thread = channel x half): relu(F + b) * relu(G + b), max over K, store
//       MS/MH/MA/MF  warps 16-19: one thread each issues the tcgen05.mma of stage 0 / the hidden stage /
//                        attention stage 1 / the last feature stage and commits their mbarriers (r02a: ONE issuing
//                        thread for the 51 small MMAs of a unit was the bottleneck, ~110 cycles per MMA)
//     Every hand-over is an mbarrier (thread arrivals from the epilogue warps, tcgen05.commit from the MMA
//     thread); rings: X0 x2, S0/H accumulator x2 (H aliases S0), activation images x3, F|G accumulator x3
//     (TMEM 2 x 64 + 3 x 128 = 512 columns), so the gather of unit u+2, the epilogues of u+1 and the final
//     epilogue of u-1 overlap the MMAs of unit u.
//   * Last stages transposed with M = 64: D^T[ch, edge]; the 64 edges [64h, 64h+64) of a unit go to TMEM lanes
//     32q+16h .. +15 (cta_group::1 M = 64 layout, tools/m64_probe.cu), F in columns [0,64), G in [64,128) of
//     the SAME lane: the product needs no shuffle and the max over a centre's K consecutive edges is a run of
//     FMNMX in one thread.
//
// Every wait is bounded (tc::mbar_wait traps) so a protocol error fails the launch instead of hanging.
#include "gridconv_tc.cuh"

#include <cstdlib>

namespace gg {

constexpr int kWsThreads = 640;
constexpr int kDX = 2;  // X0 input-image ring == S0/H accumulator ring
constexpr int kDI = 3;  // activation-image ring == F|G accumulator ring
constexpr uint32_t kPanel = 2048;  // one [128 rows x 4 k] panel of a K-major image

struct FirstWsLayout {
    int H0p, A0p, N0, H1n, H1p, KX;
    uint32_t w0_hi, w0_lo, wfh, wff_hi, wff_lo, wa1_hi, wa1_lo, bias_h, x0, img, img_stride, xf_lo, xa_hi, xa_lo, total;
};

__host__ __device__ inline FirstWsLayout first_ws_layout(const TcParams &p) {
    FirstWsLayout L;
    L.H0p = pad_to(p.f0_cout, 8);
    L.A0p = p.a1.Kp;
    L.N0 = pad_to(L.H0p + L.A0p, 16);
    L.H1n = p.fh[0].Np;
    L.H1p = p.ff.Kp;
    L.KX = max(L.H0p, L.H1p);
    uint32_t o = 0;
    L.w0_hi = o; o += (uint32_t)L.N0 * 32u;
    L.w0_lo = o; o += (uint32_t)L.N0 * 32u;
    L.wfh = o;   o += 2u * (uint32_t)p.fh[0].Np * (uint32_t)p.fh[0].Kp * 4u;
    L.wff_hi = o; o += (uint32_t)(p.ff.Kp / 4) * 1024u;
    L.wff_lo = o; o += (uint32_t)(p.ff.Kp / 4) * 1024u;
    L.wa1_hi = o; o += (uint32_t)(L.A0p / 4) * 1024u;
    L.wa1_lo = o; o += (uint32_t)(L.A0p / 4) * 1024u;
    L.bias_h = o; o += 64u * 4u;
    o = (o + 127u) & ~127u;
    L.x0 = o; o += (uint32_t)kDX * 4u * kPanel;  // per slot: hi panels 0-1, lo panels 0-1
    L.img = o;
    L.xf_lo = (uint32_t)(L.KX / 4) * kPanel;
    L.xa_hi = 2u * L.xf_lo;
    L.xa_lo = L.xa_hi + (uint32_t)(L.A0p / 4) * kPanel;
    L.img_stride = L.xa_lo + (uint32_t)(L.A0p / 4) * kPanel;
    o += (uint32_t)kDI * L.img_stride;
    L.total = o;
    return L;
}

// relu -> hi/lo split (hardware truncation model, tc_common.cuh split_op<3>) of four accumulator columns,
// one 16-byte store per image
__device__ __forceinline__ void relu_split_store4(const uint32_t *v, const float4 b, uint8_t *dst_hi, uint8_t *dst_lo) {
    float x[4], lo[4];
    x[0] = fmaxf(__uint_as_float(v[0]) + b.x, 0.f);
    x[1] = fmaxf(__uint_as_float(v[1]) + b.y, 0.f);
    x[2] = fmaxf(__uint_as_float(v[2]) + b.z, 0.f);
    x[3] = fmaxf(__uint_as_float(v[3]) + b.w, 0.f);
#pragma unroll
    for (int i = 0; i < 4; i++) lo[i] = x[i] - __uint_as_float(__float_as_uint(x[i]) & 0xFFFFE000u);
    *reinterpret_cast<float4 *>(dst_hi) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4 *>(dst_lo) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}
__device__ __forceinline__ void relu_split_store4_nobias(const uint32_t *v, uint8_t *dst_hi, uint8_t *dst_lo) {
    float x[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        x[i] = fmaxf(__uint_as_float(v[i]), 0.f);
        lo[i] = x[i] - __uint_as_float(__float_as_uint(x[i]) & 0xFFFFE000u);
    }
    *reinterpret_cast<float4 *>(dst_hi) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4 *>(dst_lo) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

template <int H0P, int A0P, int H1P>
__global__ void __launch_bounds__(kWsThreads, 1)
edge_first_ws_kernel(const __grid_constant__ TcParams p, int num_units, int cpt, int log2k) {
    constexpr int N0 = (H0P + A0P + 15) / 16 * 16;
    extern __shared__ __align__(1024) uint8_t smem[];
    // x0_full[2] s0_done[2] h_done[2] | e0_done[3] eh_done[3] f_full[3] g_full[3] fg_free[3]
    __shared__ uint64_t bars[3 * kDX + 5 * kDI];
    __shared__ uint32_t tmem_base_s;
    uint64_t *x0_full = bars, *s0_done = bars + kDX, *h_done = bars + 2 * kDX;
    uint64_t *e0_done = bars + 3 * kDX, *eh_done = e0_done + kDI, *f_full = eh_done + kDI, *g_full = f_full + kDI,
             *fg_free = g_full + kDI;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ConvParams &c = p.c;
    const int C = c.Cout;
    const FirstWsLayout L = first_ws_layout(p);

    // ---- one-time set-up: TMEM, barriers, resident weight images ----
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 32) {
        for (int i = 0; i < kDX; i++) {
            tc::mbar_init(&x0_full[i], 128);
            tc::mbar_init(&s0_done[i], 1);
            tc::mbar_init(&h_done[i], 1);
        }
        for (int i = 0; i < kDI; i++) {
            tc::mbar_init(&e0_done[i], 128);
            tc::mbar_init(&eh_done[i], 128);
            tc::mbar_init(&f_full[i], 1);
            tc::mbar_init(&g_full[i], 1);
            tc::mbar_init(&fg_free[i], 128);
        }
        tc::mbar_init_fence();
    }
    {
        // merged stage-0 image W0[N0 x 8]: rows [0, H0p) feature stage 0 = (0, w_x, w_y, w_z, 0, 0, 0, b),
        // rows [H0p, H0p + A0p) folded attention stage 0 = (w_dist, W_d + W_n | W_c + W_n, b)
        const uint32_t lbo0 = (uint32_t)L.N0 * 16u;
        for (int i = tid; i < L.N0 * 8; i += kWsThreads) {
            const int n = i >> 3, k = i & 7;
            float w = 0.f;
            if (n < L.H0p) {
                if (n < p.f0_cout) {
                    if (k >= 1 && k <= 3) w = __ldg(p.f0_w + (size_t)n * 3 + (k - 1));
                    else if (k == 7) w = __ldg(p.f0_b + n);
                }
            } else if (n < L.H0p + L.A0p) {
                w = att0_folded_weight(p.a0_w, p.a0_b, p.a0_cin, p.a0_cout, n - L.H0p, k);
            }
            float hi, lo;
            tc::split_tf32(w, hi, lo);
            const uint32_t off = tc::kmajor_off((uint32_t)n, (uint32_t)k, lbo0);
            *reinterpret_cast<float *>(smem + L.w0_hi + off) = hi;
            *reinterpret_cast<float *>(smem + L.w0_lo + off) = lo;
        }
        {   // hidden stage: packed plain image (hi then lo), as is
            const TcStage &st = p.fh[0];
            const int n4 = 2 * st.Np * st.Kp / 4;
            const float4 *src = reinterpret_cast<const float4 *>(p.packed + st.w_off);
            float4 *dst = reinterpret_cast<float4 *>(smem + L.wfh);
            for (int i = tid; i < n4; i += kWsThreads) dst[i] = __ldg(src + i);
        }
        // rows 0..63 of every [128 x 4] panel of the packed transposed-stage images (chunk 0)
        auto load_compact = [&](const TcStage &st, uint32_t dst_hi, uint32_t dst_lo) {
            const int per_img = (st.Kp / 4) * 256;
            for (int i = tid; i < 2 * per_img; i += kWsThreads) {
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
        for (int i = tid; i < 64; i += kWsThreads)
            reinterpret_cast<float *>(smem + L.bias_h)[i] = i < p.fh[0].Cout ? __ldg(p.fh[0].bias + i) : 0.f;
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    const int n_my = (int)blockIdx.x < num_units ? (num_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const unsigned centers_total = (unsigned)c.B * (unsigned)c.O;
    const int out_w = 4 + C;
    const uint32_t q = (uint32_t)(warp & 3);
    const uint32_t row = q * 32u + (uint32_t)lane;                          // edge row of the unit (G, E0, EH)
    const uint32_t row_off = (row >> 3) * 128u + (row & 7u) * 16u;           // its offset inside a panel
    const uint32_t lane_base = (q * 32u) << 16;                             // TMEM lane quadrant of this warp

    if (warp < 4) {
        // =========================== G: gather + input image ===========================
        const int r = tid;
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
            if (i >= kDX) tc::mbar_wait(&s0_done[x], (uint32_t)((i - kDX) >> 1) & 1u);  // MMAs of unit i-2 read this slot
            uint8_t *x0 = smem + L.x0 + (uint32_t)x * 4u * kPanel + row_off;
            *reinterpret_cast<float4 *>(x0) = make_float4(in[0], in[1], in[2], in[3]);
            *reinterpret_cast<float4 *>(x0 + kPanel) = make_float4(in[4], in[5], in[6], in[7]);
            *reinterpret_cast<float4 *>(x0 + 2 * kPanel) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4 *>(x0 + 3 * kPanel) = make_float4(lo[4], lo[5], lo[6], lo[7]);
            tc::fence_async_smem();
            tc::mbar_arrive(&x0_full[x]);
            // centre columns of the output rows
            if (r < cpt * 4) {
                const unsigned center = (unsigned)(blockIdx.x + i * gridDim.x) * (unsigned)cpt + (unsigned)(r >> 2);
                if (center < centers_total)
                    c.out[(size_t)center * out_w + (r & 3)] = __ldg(reinterpret_cast<const float *>(c.cent + center) + (r & 3));
            }
            head = head_n; cent = cent_n; v_cur = v_nxt;
            idx_nxt = idx_n2; v_nxt = v_n2;
        }
    } else if (warp < 8) {
        // =========================== E0: stage-0 epilogue ===========================
        int d = 0;
        uint32_t ph_d = 0;  // use count parity of image slot d (== F|G slot d)
        for (int i = 0; i < n_my; i++) {
            const int a = i & 1;
            if (i >= kDI) {  // last-stage MMAs of unit i-3 have read the images
                tc::mbar_wait(&f_full[d], ph_d ^ 1u);
                tc::mbar_wait(&g_full[d], ph_d ^ 1u);
            }
            tc::mbar_wait(&s0_done[a], (uint32_t)(i >> 1) & 1u);
            tc::fence_after_sync();
            const uint32_t taddr = tmem + lane_base + (uint32_t)a * 64u;
            uint8_t *img = smem + L.img + (uint32_t)d * L.img_stride + row_off;
            uint32_t v[N0];
#pragma unroll
            for (int c0 = 0; c0 < N0; c0 += 16) tc::tmem_ld16(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(v + c0));
            tc::tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < H0P; cc += 4) {
                uint8_t *dst = img + (uint32_t)(cc >> 2) * kPanel;
                relu_split_store4_nobias(v + cc, dst, dst + L.xf_lo);
            }
#pragma unroll
            for (int cc = 0; cc < A0P; cc += 4) {
                uint8_t *dst = img + L.xa_hi + (uint32_t)(cc >> 2) * kPanel;
                relu_split_store4_nobias(v + H0P + cc, dst, dst + (L.xa_lo - L.xa_hi));
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::mbar_arrive(&e0_done[d]);
            if (++d == kDI) { d = 0; ph_d ^= 1u; }
        }
    } else if (warp < 12) {
        // =========================== EH: hidden-stage epilogue ===========================
        const float *bias_h = reinterpret_cast<const float *>(smem + L.bias_h);
        int d = 0;
        for (int i = 0; i < n_my; i++) {
            const int a = i & 1;
            tc::mbar_wait(&h_done[a], (uint32_t)(i >> 1) & 1u);
            tc::fence_after_sync();
            const uint32_t taddr = tmem + lane_base + (uint32_t)a * 64u;
            uint8_t *img = smem + L.img + (uint32_t)d * L.img_stride + row_off;
            constexpr int H1L = (H1P + 15) / 16 * 16;
            uint32_t v[H1L];
#pragma unroll
            for (int c0 = 0; c0 < H1L; c0 += 16) tc::tmem_ld16(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(v + c0));
            tc::tmem_ld_wait();
#pragma unroll
            for (int cc = 0; cc < H1P; cc += 4) {
                uint8_t *dst = img + (uint32_t)(cc >> 2) * kPanel;
                relu_split_store4(v + cc, *reinterpret_cast<const float4 *>(bias_h + cc), dst, dst + L.xf_lo);
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::mbar_arrive(&eh_done[d]);
            if (++d == kDI) d = 0;
        }
    } else if (warp < 16) {
        // =========================== EF: final epilogue ===========================
        // lane l of warp q: channel 16q + (l & 15), edges [64h, 64h + 64) of the unit with h = l >> 4.
        // (Tried, r02: starting the accumulators at the bias through tcgen05.st.x16 -- the 16 source registers of
        // every store have to be filled with MOVs, which costs what the bias adds cost.)
        const int h = lane >> 4;
        const int ch = 16 * (int)q + (lane & 15);
        const bool chv = ch < C;
        const float bf = chv ? __ldg(p.ff.bias + ch) : 0.f;
        const float bg = chv ? __ldg(p.a1.bias + ch) : 0.f;
        const float pre_floor = c.pre_relu ? 0.f : -3.402823466e+38f;
        float *out_ch = c.out + 4 + ch;
        const int kmask = (1 << log2k) - 1;
        int f = 0;
        uint32_t ph_f = 0;
        for (int i = 0; i < n_my; i++) {
            const unsigned c_base = (unsigned)(blockIdx.x + i * gridDim.x) * (unsigned)cpt;
            tc::mbar_wait(&f_full[f], ph_f);
            tc::mbar_wait(&g_full[f], ph_f);
            tc::fence_after_sync();
            const uint32_t taddr = tmem + lane_base + 128u + (uint32_t)f * 128u;
            float m = -3.402823466e+38f;
            uint32_t fa[16], ga[16], fb[16], gb[16];
            auto reduce16 = [&](const uint32_t (&fv)[16], const uint32_t (&gv)[16], int c0) {
                float pr[16];
#pragma unroll
                for (int j = 0; j < 16; j++)
                    pr[j] = fmaxf(__uint_as_float(fv[j]) + bf, 0.f) * fmaxf(__uint_as_float(gv[j]) + bg, 0.f);  // :167 att * feats
                float mm = pr[0];
#pragma unroll
                for (int j = 1; j < 16; j++) mm = fmaxf(mm, pr[j]);
                m = fmaxf(m, mm);
                const int e_end = 64 * h + c0 + 16;  // edges of this lane reduced so far end here
                if (log2k <= 6 && (e_end & kmask) == 0) {
                    const unsigned center = c_base + (unsigned)((e_end >> log2k) - 1);
                    if (chv && center < centers_total)
                        out_ch[(size_t)center * out_w] = fmaxf(m, pre_floor) * __ldg(c.centmsk + center);
                    m = -3.402823466e+38f;
                }
            };
            tc::tmem_ld16(taddr, fa);
            tc::tmem_ld16(taddr + 64u, ga);
            tc::tmem_ld_wait();
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                tc::tmem_ld16(taddr + (uint32_t)(c0 + 16), fb);
                tc::tmem_ld16(taddr + (uint32_t)(c0 + 80), gb);
                reduce16(fa, ga, c0);
                tc::tmem_ld_wait();
                if (c0 == 0) {
                    tc::tmem_ld16(taddr + 32u, fa);
                    tc::tmem_ld16(taddr + 96u, ga);
                }
                reduce16(fb, gb, c0 + 16);
                tc::tmem_ld_wait();
            }
            // every TMEM read of this unit is done: the slot may be overwritten
            tc::fence_before_sync();
            tc::mbar_arrive(&fg_free[f]);
            if (log2k == 7) {  // one centre per unit: its two 64-edge halves meet across the lane pair
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
                if (h == 0 && chv && c_base < centers_total)
                    out_ch[(size_t)c_base * out_w] = fmaxf(m, pre_floor) * __ldg(c.centmsk + c_base);
            }
            if (++f == kDI) { f = 0; ph_f ^= 1u; }
        }
    } else if (lane == 0) {
        // =========================== MMA issue: one thread per stage ===========================
        // Descriptors are built once; per use only the 16-byte-unit start-address field (bits 0-13, never
        // overflowing: shared addresses < 256 KB) is advanced by an addition.
        const uint32_t sb = tc::smem_u32(smem);
        auto adv = [](uint64_t dsc, uint32_t bytes) -> uint64_t { return dsc + (uint64_t)(bytes >> 4); };
        if (warp == 16) {
            // ---- MS: stage 0, D0[edge, H0p + A0p], K = 8.  The accumulator slot held the hidden accumulator
            //      of unit i-2: wait for its epilogue ----
            const uint32_t idesc = tc::make_idesc_tf32(128, N0);
            const uint32_t lbo_w0 = (uint32_t)N0 * 16u;
            const uint64_t w0h = tc::make_sdesc(sb + L.w0_hi, lbo_w0), w0l = tc::make_sdesc(sb + L.w0_lo, lbo_w0);
            const uint64_t xh0 = tc::make_sdesc(sb + L.x0, kPanel), xl0 = tc::make_sdesc(sb + L.x0 + 2u * kPanel, kPanel);
            int d2 = 0;
            uint32_t ph2 = 0;  // slot / parity of unit i-2 in the image ring
            for (int i = 0; i < n_my; i++) {
                const uint32_t x = (uint32_t)(i & 1);
                tc::mbar_wait(&x0_full[x], (uint32_t)(i >> 1) & 1u);
                if (i >= kDX) {
                    tc::mbar_wait(&eh_done[d2], ph2);
                    if (++d2 == kDI) { d2 = 0; ph2 ^= 1u; }
                }
                tc::fence_after_sync();
                const uint64_t ah = adv(xh0, x * 4u * kPanel), al = adv(xl0, x * 4u * kPanel);
                const uint32_t dacc = tmem + x * 64u;
                tc::mma_tf32(dacc, al, w0h, idesc, 0);
                tc::mma_tf32(dacc, ah, w0l, idesc, 1);
                tc::mma_tf32(dacc, ah, w0h, idesc, 1);
                tc::mma_commit(&s0_done[x]);
            }
        } else if (warp == 17) {
            // ---- MH: hidden feature stage D[edge, ch] (aliases the stage-0 accumulator) ----
            const uint32_t idesc = tc::make_idesc_tf32(128, L.H1n);
            const uint32_t lbo_wh = (uint32_t)L.H1n * 16u;
            const uint64_t wh0 = tc::make_sdesc(sb + L.wfh, lbo_wh);
            const uint64_t wl0 = tc::make_sdesc(sb + L.wfh + (uint32_t)L.H1n * (uint32_t)L.H0p * 4u, lbo_wh);
            const uint64_t xh0 = tc::make_sdesc(sb + L.img, kPanel), xl0 = tc::make_sdesc(sb + L.img + L.xf_lo, kPanel);
            constexpr int ks_h = H0P / 8;
            int d = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_my; i++) {
                tc::mbar_wait(&e0_done[d], ph);
                tc::fence_after_sync();
                uint64_t ah = adv(xh0, (uint32_t)d * L.img_stride), al = adv(xl0, (uint32_t)d * L.img_stride);
                uint64_t bh = wh0, bl = wl0;
                const uint32_t dacc = tmem + (uint32_t)(i & 1) * 64u;
                uint32_t acc = 0;
                for (int ks = 0; ks < ks_h; ks++) {
                    tc::mma_tf32(dacc, al, bh, idesc, acc);
                    tc::mma_tf32(dacc, ah, bl, idesc, 1);
                    tc::mma_tf32(dacc, ah, bh, idesc, 1);
                    acc = 1;
                    ah = adv(ah, 2u * kPanel); al = adv(al, 2u * kPanel);
                    bh = adv(bh, 2u * lbo_wh); bl = adv(bl, 2u * lbo_wh);
                }
                tc::mma_commit(&h_done[i & 1]);
                if (++d == kDI) { d = 0; ph ^= 1u; }
            }
        } else if (warp == 18 || warp == 19) {
            // ---- MA (warp 18): attention stage 1 -> G;  MF (warp 19): last feature stage -> F.  Transposed,
            //      M = 64 channels x N = 64 edges per half; the two halves' MMAs alternate (independent accumulators) ----
            const bool is_f = warp == 19;
            const uint32_t idesc = tc::make_idesc_tf32(64, 64);
            const uint64_t wh0 = tc::make_sdesc(sb + (is_f ? L.wff_hi : L.wa1_hi), 1024u);
            const uint64_t wl0 = tc::make_sdesc(sb + (is_f ? L.wff_lo : L.wa1_lo), 1024u);
            const uint64_t xh0 = tc::make_sdesc(sb + L.img + (is_f ? 0u : L.xa_hi), kPanel);
            const uint64_t xl0 = tc::make_sdesc(sb + L.img + (is_f ? L.xf_lo : L.xa_lo), kPanel);
            const int nks = (is_f ? H1P : A0P) / 8;
            uint64_t *ready = is_f ? eh_done : e0_done, *full = is_f ? f_full : g_full;
            const uint32_t col0 = 128u + (is_f ? 0u : 64u);
            int d = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_my; i++) {
                tc::mbar_wait(&ready[d], ph);
                if (i >= kDI) tc::mbar_wait(&fg_free[d], ph ^ 1u);
                tc::fence_after_sync();
                uint64_t ah = wh0, al = wl0;
                uint64_t bh = adv(xh0, (uint32_t)d * L.img_stride), bl = adv(xl0, (uint32_t)d * L.img_stride);
                const uint32_t dc0 = tmem + col0 + (uint32_t)d * 128u, dc1 = dc0 + (16u << 16);
                uint32_t acc = 0;
                for (int ks = 0; ks < nks; ks++) {
                    const uint64_t bh1 = adv(bh, 1024u), bl1 = adv(bl, 1024u);  // rows 64..127 of the image
                    tc::mma_tf32(dc0, al, bh, idesc, acc);
                    tc::mma_tf32(dc1, al, bh1, idesc, acc);
                    acc = 1;
                    tc::mma_tf32(dc0, ah, bl, idesc, 1);
                    tc::mma_tf32(dc1, ah, bl1, idesc, 1);
                    tc::mma_tf32(dc0, ah, bh, idesc, 1);
                    tc::mma_tf32(dc1, ah, bh1, idesc, 1);
                    ah = adv(ah, 2048u); al = adv(al, 2048u);
                    bh = adv(bh, 2u * kPanel); bl = adv(bl, 2u * kPanel);
                }
                tc::mma_commit(&full[d]);
                if (++d == kDI) { d = 0; ph ^= 1u; }
            }
        }
    }
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
    if (p.nsplit != 3 || !p.has_ff || !p.has_att || !p.f0_cuda || p.nfh != 1 || p.dbg != nullptr) return -1;
    if (!(c.K == 16 || c.K == 32 || c.K == 64 || c.K == 128)) return -1;
    if (c.Cout > 64 || p.ff.Np != 128 || p.a1.Np != 128) return -1;
    const FirstWsLayout L = first_ws_layout(p);
    if (L.N0 > 64 || L.H1n > 64 || L.H1p > 64 || p.fh[0].Kp != L.H0p || p.fh[0].Cout > 64) return -1;
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
