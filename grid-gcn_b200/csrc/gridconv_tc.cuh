// Shared types and device helpers of the tensor-core GridConv kernels (gridconv_tc.cu, gridconv_first_ws.cu).
#pragma once
#include "gridconv_common.cuh"
#include "tc_common.cuh"

namespace gg {

constexpr int kTcThreads = 160;     // warps 0-3: gather + TMEM epilogue (thread = TMEM lane); warp 4: TMA + MMA issue
constexpr int kTileRows = 128;
#ifndef GG_SLICE_K
#define GG_SLICE_K 32   // k extent of one streamed weight slice (tools/build_variant.py -DGG_SLICE_K=16)
#endif
#ifndef GG_WIDE_ALWAYS
#define GG_WIDE_ALWAYS 0  // 1: the 8-warp wide kernel also where three 5-warp CTAs per SM would fit (layer 1)
#endif
#ifndef GG_RING_CAP
#define GG_RING_CAP 4   // most ring slots the row-major kernel A / kernel B take (transposed kernel A: + 2)
#endif
constexpr int kSliceK = GG_SLICE_K;  // k extent of one streamed weight slice
constexpr int kSlotBytes = 2 * 128 * kSliceK * 4;  // hi + lo images of a [128 x 32] slice
constexpr int kMaxRing = 8;
constexpr int kRowsThreads = 288;    // row-major kernel A: warps 0-7 workers (two per TMEM lane quadrant), warp 8 TMA + MMA

struct TcStage {
    int transposed;  // 0: D[row, ch], weights resident in smem as the B operand; 1: D^T[ch, row], weights streamed as A
    int Cin, Cout;   // logical dims
    int Kp, Np;      // Cin padded to 8; Cout padded to 16 (plain) / 128 (transposed)
    long long w_off; // float offset of this stage inside the packed buffer
    const float *bias;
};

struct TcParams {
    ConvParams c;
    const float *packed;  // packed hi/lo operand images of every tensor-core stage
    float *ftab;          // kernel A output / kernel B input: (B*Nprev, Cout) transformed features
    int nsplit;           // 1 or 3
    // kernel A: stages a[0..na)
    int na;
    TcStage a[GRIDGCN_MAX_STAGES];
    int a_rows;           // rows per tile (64 or 128)
    // kernel B.  Stages whose K dimension is tiny run on the CUDA cores inside the gather phase
    // (exact fp32 FMA): the first-layer feature stage 0 (K = 3) and the attention stage 0 (K <= 10).
    int f0_cuda, f0_cout;
    const float *f0_w, *f0_b;      // (f0_cout, 3), (f0_cout)
    int a0_cin, a0_cout;
    const float *a0_w, *a0_b;      // (a0_cout, a0_cin), (a0_cout)
    int nfh;                       // remaining first-layer hidden feature stages (tensor core, plain)
    TcStage fh[GRIDGCN_MAX_STAGES];
    int has_ff;                    // first layer: last feature stage (tensor core, transposed)
    TcStage ff;
    int has_att;                   // attention stage 1 (tensor core, transposed)
    TcStage a1;
    int ring_slots, ring_sticky;   // weight ring: number of slots; sticky = whole per-tile sequence resident
    int tmem_cols;
    unsigned long long *dbg;       // optional phase-cycle counters (debug, see gridgcn_debug_phase_buffer)
};

__host__ __device__ inline int pad_to(int x, int m) { return (x + m - 1) / m * m; }

// Bounded mbarrier wait (hardware sleep between tries, a few seconds in total), then trap: never hang the GPU.
__device__ __forceinline__ void wait_bar(uint64_t *bar, uint32_t parity) { tc::mbar_wait(bar, parity); }

// ---- attention stage 0, folded --------------------------------------------------------------------
// att_vec = [dist, d, c, n] with n = c + d (gcn_module_g_att.py:209-222), so
//     W a + b = w_dist dist + (W_d + W_n) d + (W_c + W_n) c + b :
// 7 multiply-adds per channel instead of 10 and two 16-byte weight loads instead of three.  The kernels
// keep the folded weights in shared memory as wa0_s[kh][8] = (w_dist, wd_x, wd_y, wd_z | wc_x, wc_y, wc_z, b).
struct Att7 {
    float dist, dx, dy, dz, cx, cy, cz;
};
__device__ __forceinline__ Att7 fold_att(int attfdim, const float *att) {
    Att7 a;
    if (attfdim <= 3) {
        a.dist = 0.f; a.dx = att[0]; a.dy = att[1]; a.dz = att[2];
    } else {
        a.dist = att[0]; a.dx = att[1]; a.dy = att[2]; a.dz = att[3];
    }
    const bool full = attfdim >= 10;
    a.cx = full ? att[4] : 0.f; a.cy = full ? att[5] : 0.f; a.cz = full ? att[6] : 0.f;
    return a;
}
// element q (0..7) of channel j's folded row
__device__ __forceinline__ float att0_folded_weight(const float *w, const float *b, int cin, int cout, int j,
                                                   int q) {
    if (j >= cout) return 0.f;
    const float *r = w + (size_t)j * cin;
    if (q == 7) return __ldg(b + j);
    if (cin >= 10) {
        if (q == 0) return __ldg(r);
        if (q < 4) return __ldg(r + q) + __ldg(r + q + 6);        // W_d + W_n
        return __ldg(r + q) + __ldg(r + q + 3);                   // W_c + W_n   (q = 4..6)
    }
    if (q >= 4) return 0.f;
    if (cin == 4) return __ldg(r + q);
    return q == 0 ? 0.f : (q - 1 < cin ? __ldg(r + q - 1) : 0.f);  // [d] only
}
__device__ __forceinline__ float att0_channel(const float4 wd, const float4 wc, const Att7 &a) {
    float acc = wc.w;
    acc = fmaf(wc.x, a.cx, acc); acc = fmaf(wc.y, a.cy, acc); acc = fmaf(wc.z, a.cz, acc);
    acc = fmaf(wd.x, a.dist, acc); acc = fmaf(wd.y, a.dx, acc);
    acc = fmaf(wd.z, a.dy, acc); acc = fmaf(wd.w, a.dz, acc);
    return acc;
}

}  // namespace gg
