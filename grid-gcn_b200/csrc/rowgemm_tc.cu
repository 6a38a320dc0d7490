// Row GEMM stage on the tensor cores (tcgen05 kind::tf32, 3-pass hi/lo split = fp32-class accuracy):
//     out[r, 0:N] = act_out( W * act_in([in1[r, :] | in2[r, :]]) + b ) * scale[r]
// the same operator as rowmlp.cu (decoder centre branch / update MLP / heads: reference
// segmentation/models/gcn_module_g_att.py:267-285, :24-43, ggcn_models_g.py:30-36; classification head
// classification/models/ggcn_models_g.py:25-35) and the per-point feature MLP stages of the hoisted GridConv
// layers (gridconv_tc.cu kernel A) -- r02 launch list of the shipped ScanNet graph at B = 12: the CUDA-core row
// MLP was 60 % of the step (1.2 of 2.0 ms, 11 TFLOP/s).
//
// One persistent, warp-specialised CTA per SM, tiles of 128 rows, K in chunks of 32:
//   loaders  warps 0-7 (two groups of four, alternating chunks; thread = row): 128-bit global loads of the
//            row's 32 inputs (and of one row of the weight chunk), act_in, hi/lo split, K-major operand images
//            into a shared-memory ring (conflict-free 16-byte stores);
//   MMA      warp 12, one thread: D[row, n] += X_chunk * W_chunk^T, 4 k-steps x 3 passes per chunk, accumulator in
//            TMEM (double buffered: the epilogue of tile t overlaps the MMAs of tile t+1); tcgen05.commit frees
//            the ring slot;
//   epilogue warps 8-11 (thread = row): TMEM -> +bias, act_out, row scale -> 128-bit stores of the output row.
// Weights are read raw (N x K row-major fp32): no packing step, every CTA re-reads the same chunk from L2.
#include "../../include/gridgcn_b200.h"
#include "common.cuh"
#include "tc_common.cuh"

#include <algorithm>

namespace gg {

constexpr int kRgThreads = 416;  // 8 loader warps, 4 epilogue warps, 1 MMA warp
constexpr int kRgKC = 32;        // k extent of one ring slot
constexpr uint32_t kRgPanel = 2048;

struct RowGemmParams {
    const float *in1, *in2, *W, *bias, *scale, *head4;
    float *out, *out_head;
    int ld1, c1, ld2, c2, N, Np, K, nchunk, relu_in, relu_out, ldo, stages;
    long long rows;
    int tiles;
};

template <bool WIDE>  // WIDE: 128 < N <= 256 (two weight rows per loader thread, 256-column accumulators)
__global__ void __launch_bounds__(kRgThreads, 1) rowgemm_ws_kernel(const __grid_constant__ RowGemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[2 * 4 + 4];
    __shared__ uint32_t tmem_base_s;
    __shared__ float bias_s[256];
    uint64_t *full = bars, *empty = bars + 4, *acc_full = bars + 8, *acc_free = bars + 10;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Np = p.Np, S = p.stages;
    const uint32_t lbo_w = (uint32_t)Np * 16u;
    const uint32_t w_img = (uint32_t)Np * 128u;                 // one [Np x 32] image
    const uint32_t stage_bytes = 2u * 8u * kRgPanel + 2u * w_img;  // X hi | X lo | W hi | W lo
    constexpr uint32_t acc_cols = WIDE ? 256u : 128u;

    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 2 * acc_cols <= 256 ? 256 : 512);
    if (tid == 32) {
        for (int i = 0; i < S; i++) {
            tc::mbar_init(&full[i], 4);   // one arrival per warp of the loader group
            tc::mbar_init(&empty[i], 1);  // tcgen05.commit
        }
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(&acc_full[i], 1);
            tc::mbar_init(&acc_free[i], 4);
        }
        tc::mbar_init_fence();
    }
    for (int i = tid; i < 256; i += kRgThreads) bias_s[i] = i < p.N ? __ldg(p.bias + i) : 0.f;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const int n_my = (int)blockIdx.x < p.tiles ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nchunk = p.nchunk;

    if (warp < 8) {
        // =========================== loaders ===========================
        // Global loads are COALESCED (8 lanes read the 128 contiguous bytes of a row's chunk, 4 rows per instruction;
        // r02l: one row per thread meant 32 sectors per request and an LSU-bound kernel at 1.7 TB/s) and transposed to
        // the thread = row arrangement of the operand images through a warp-private staging tile.
        const int grp = warp >> 2;                     // running chunks `it` with (it & 1) == grp
        const int wq = warp & 3;                       // this warp's 32 tile rows / weight rows
        float *stage = reinterpret_cast<float *>(smem + (size_t)S * stage_bytes) + warp * (16 * 36);
        const int lrow = lane >> 3, lg = lane & 7;     // load arrangement: row 4j + lrow, group lg
        const int prow = lane & 15, pg0 = 4 * (lane >> 4);  // processing arrangement: row 16h + prow, groups pg0..pg0+3
        long long it = 0;                              // running chunk number of this CTA (ring position)
        for (int i = 0; i < n_my; i++) {
            const long long row0 = (long long)(blockIdx.x + i * gridDim.x) * 128 + 32 * wq;
            for (int c = 0; c < nchunk; c++, it++) {
                if ((int)(it & 1) != grp) continue;  // alternate on the RUNNING chunk number: both groups work when nchunk is odd
                const int slot = (int)(it % S);
                const uint32_t use = (uint32_t)(it / S);
                const int k = c * kRgKC + 4 * lg;
                float4 xin[8], win[8], win2[WIDE ? 8 : 1];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const long long row = row0 + 4 * j + lrow;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < p.rows && k < p.K)
                        v = k < p.c1 ? __ldg(reinterpret_cast<const float4 *>(p.in1 + row * p.ld1 + k))
                                     : __ldg(reinterpret_cast<const float4 *>(p.in2 + row * p.ld2 + (k - p.c1)));
                    xin[j] = v;
                    const int n = 32 * wq + 4 * j + lrow;
                    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (n < p.N && k < p.K) w = __ldg(reinterpret_cast<const float4 *>(p.W + (size_t)n * p.K + k));
                    win[j] = w;
                    if (WIDE) {
                        float4 w2 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (n + 128 < p.N && k < p.K) w2 = __ldg(reinterpret_cast<const float4 *>(p.W + (size_t)(n + 128) * p.K + k));
                        win2[j] = w2;
                    }
                }
                // in[j] = (row 4j + lrow, group lg)  ->  out[4h + gg] = (row 16h + prow, group pg0 + gg)
                auto transpose = [&](const float4 *in, float4 *out) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            *reinterpret_cast<float4 *>(stage + (4 * j + lrow) * 36 + 4 * lg) = in[4 * h + j];
                        __syncwarp();
#pragma unroll
                        for (int gg = 0; gg < 4; gg++)
                            out[4 * h + gg] = *reinterpret_cast<const float4 *>(stage + prow * 36 + 4 * (pg0 + gg));
                        __syncwarp();
                    }
                };
                float4 xv[8], wv[8], wv2[WIDE ? 8 : 1];
                transpose(xin, xv);
                transpose(win, wv);
                if (WIDE) transpose(win2, wv2);
                if (use > 0) tc::mbar_wait(&empty[slot], (use - 1) & 1u);  // the MMAs that read this slot have retired
                uint8_t *st = smem + (size_t)slot * stage_bytes;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t R = (uint32_t)(32 * wq + 16 * h + prow);  // tile row == weight row
                    const uint32_t row_off = (R >> 3) * 128u + (R & 7u) * 16u;
                    uint8_t *xh = st + row_off, *xl = xh + 8u * kRgPanel;
                    uint8_t *wh = st + 16u * kRgPanel + row_off, *wl = wh + w_img;
#pragma unroll
                    for (int gg = 0; gg < 4; gg++) {
                        const uint32_t g = (uint32_t)(pg0 + gg);
                        float4 v = xv[4 * h + gg];
                        if (p.relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        float4 lo;  // activations: hardware truncation model (tc_common.cuh split_op<3>)
                        lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                        lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                        lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                        lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                        *reinterpret_cast<float4 *>(xh + g * kRgPanel) = v;
                        *reinterpret_cast<float4 *>(xl + g * kRgPanel) = lo;
                        if ((int)R < Np) {
                            float4 hh, l;  // weights: round-to-nearest parts, as the packed images of gridconv_tc.cu
                            const float4 w = wv[4 * h + gg];
                            tc::split_tf32(w.x, hh.x, l.x); tc::split_tf32(w.y, hh.y, l.y);
                            tc::split_tf32(w.z, hh.z, l.z); tc::split_tf32(w.w, hh.w, l.w);
                            *reinterpret_cast<float4 *>(wh + g * lbo_w) = hh;
                            *reinterpret_cast<float4 *>(wl + g * lbo_w) = l;
                        }
                        if (WIDE) {
                            float4 hh, l;
                            const float4 w = wv2[4 * h + gg];
                            tc::split_tf32(w.x, hh.x, l.x); tc::split_tf32(w.y, hh.y, l.y);
                            tc::split_tf32(w.z, hh.z, l.z); tc::split_tf32(w.w, hh.w, l.w);
                            *reinterpret_cast<float4 *>(wh + 2048u + g * lbo_w) = hh;  // rows 128.. : 16 row groups further
                            *reinterpret_cast<float4 *>(wl + 2048u + g * lbo_w) = l;
                        }
                    }
                }
                tc::fence_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&full[slot]);
            }
        }
    } else if (warp < 12) {
        // =========================== epilogue ===========================
        // thread = row out of TMEM; the output rows leave through a warp-private staging tile so that 4 lanes write
        // the 64 contiguous bytes of a row's 16 columns (8 rows per store instruction, full sectors)
        const uint32_t q = (uint32_t)(warp & 3);
        const int r = (int)q * 32 + lane;
        const int N = p.N;
        float *stg = reinterpret_cast<float *>(smem + (size_t)S * stage_bytes) + 8 * (16 * 36) + (warp - 8) * (32 * 20);
        const int srow = lane >> 2, sg = lane & 3;  // store arrangement: row 8i + srow, columns 4 sg .. 4 sg + 3
        for (int i = 0; i < n_my; i++) {
            const int a = i & 1;
            const long long row_base = (long long)(blockIdx.x + i * gridDim.x) * 128 + 32 * (int)q;
            const long long row = row_base + lane;
            const bool rv = row < p.rows;
            const float s = (rv && p.scale) ? __ldg(p.scale + row) : 1.f;
            float4 head = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rv && p.head4) head = __ldg(reinterpret_cast<const float4 *>(p.head4) + row);
            tc::mbar_wait(&acc_full[a], (uint32_t)(i >> 1) & 1u);
            tc::fence_after_sync();
            const uint32_t taddr = tmem + ((q * 32u) << 16) + (uint32_t)a * acc_cols;
            for (int c0 = 0; c0 < N; c0 += 16) {
                uint32_t v[16];
                tc::tmem_ld16(taddr + (uint32_t)c0, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float x = __uint_as_float(v[4 * g + j]) + bias_s[c0 + 4 * g + j];
                        if (p.relu_out) x = fmaxf(x, 0.f);
                        o[j] = x * s;
                    }
                    *reinterpret_cast<float4 *>(stg + lane * 20 + 4 * g) = make_float4(o[0], o[1], o[2], o[3]);
                }
                __syncwarp();
                const int cc = c0 + 4 * sg;
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    const int rl = 8 * ii + srow;
                    const long long orow = row_base + rl;
                    const float4 o = *reinterpret_cast<const float4 *>(stg + rl * 20 + 4 * sg);
                    if (orow < p.rows) {
                        float *dst = p.out + orow * p.ldo + cc;
                        if (cc + 3 < N) *reinterpret_cast<float4 *>(dst) = o;
                        else {
                            if (cc < N) dst[0] = o.x;
                            if (cc + 1 < N) dst[1] = o.y;
                            if (cc + 2 < N) dst[2] = o.z;
                        }
                    }
                }
                __syncwarp();
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_free[a]);
            if (rv && p.head4) *reinterpret_cast<float4 *>(p.out_head + row * p.ldo) = head;
        }
    } else if (lane == 0) {
        // =========================== MMA issue ===========================
        const uint32_t sb = tc::smem_u32(smem);
        const uint32_t idesc = tc::make_idesc_tf32(128, Np);
        auto adv = [](uint64_t dsc, uint32_t bytes) -> uint64_t { return dsc + (uint64_t)(bytes >> 4); };
        const uint64_t xh0 = tc::make_sdesc(sb, kRgPanel), xl0 = tc::make_sdesc(sb + 8u * kRgPanel, kRgPanel);
        const uint64_t wh0 = tc::make_sdesc(sb + 16u * kRgPanel, lbo_w), wl0 = tc::make_sdesc(sb + 16u * kRgPanel + w_img, lbo_w);
        long long it = 0;
        for (int i = 0; i < n_my; i++) {
            const int a = i & 1;
            if (i >= 2) tc::mbar_wait(&acc_free[a], (uint32_t)((i >> 1) - 1) & 1u);
            tc::fence_after_sync();
            const uint32_t dacc = tmem + (uint32_t)a * acc_cols;
            for (int c = 0; c < nchunk; c++, it++) {
                const int slot = (int)(it % S);
                tc::mbar_wait(&full[slot], (uint32_t)(it / S) & 1u);
                tc::fence_after_sync();
                const uint32_t so = (uint32_t)slot * stage_bytes;
                uint64_t ah = adv(xh0, so), al = adv(xl0, so), bh = adv(wh0, so), bl = adv(wl0, so);
#pragma unroll
                for (int ks = 0; ks < kRgKC / 8; ks++) {
                    tc::mma_tf32(dacc, al, bh, idesc, (c | ks) != 0);
                    tc::mma_tf32(dacc, ah, bl, idesc, 1);
                    tc::mma_tf32(dacc, ah, bh, idesc, 1);
                    ah = adv(ah, 2u * kRgPanel); al = adv(al, 2u * kRgPanel);
                    bh = adv(bh, 2u * lbo_w); bl = adv(bl, 2u * lbo_w);
                }
                tc::mma_commit(&empty[slot]);
            }
            tc::mma_commit(&acc_full[a]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (warp == 0) tc::tmem_dealloc(tmem, 2 * acc_cols <= 256 ? 256 : 512);
}

// Returns -1 when the arguments do not fit the tensor-core kernel (alignment, widths): the caller uses the
// CUDA-core kernel of rowmlp.cu instead; otherwise the CUDA error code of the launch.
int launch_rowgemm_tc(const float *in1, int ld1, int c1, const float *in2, int ld2, int c2, const float *W,
                      const float *bias, int N, int relu_in, int relu_out, const float *scale, float *out, int ldo,
                      const float *cent, float *out_table, long long rows, cudaStream_t st) {
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const int K = c1 + c2;
    if (N < 1 || N > 256 || K < 4 || (c1 & 3) || (c2 & 3) || (ld1 & 3) || (c2 > 0 && (ld2 & 3)) || (ldo & 3)) return -1;
    if (!al16(in1) || (c2 > 0 && !al16(in2)) || !al16(W) || !al16(out) || (K & 3)) return -1;
    if (rows <= 0) return 0;
    RowGemmParams p{};
    p.in1 = in1; p.in2 = c2 > 0 ? in2 : nullptr; p.W = W; p.bias = bias; p.scale = scale; p.head4 = cent;
    p.out = out; p.out_head = out_table;
    p.ld1 = ld1; p.c1 = c1; p.ld2 = ld2; p.c2 = c2; p.N = N; p.K = K;
    p.Np = N <= 128 ? std::max(16, (N + 15) / 16 * 16) : 256;
    p.nchunk = (K + kRgKC - 1) / kRgKC;
    p.relu_in = relu_in; p.relu_out = relu_out; p.ldo = ldo; p.rows = rows;
    const long long tiles = (rows + 127) / 128;
    if (tiles > 0x7fffffff) return -1;
    p.tiles = (int)tiles;
    const size_t stage = 2 * 8 * kRgPanel + 2 * (size_t)p.Np * 128;
    const size_t staging = 8 * (16 * 36) * 4 + 4 * (32 * 20) * 4;  // warp-private transposition tiles (loaders, epilogue)
    p.stages = (int)std::min<size_t>(4, (225 * 1024 - staging) / stage);
    if (p.stages < 2) return -1;
    const size_t smem = stage * p.stages + staging;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static PerDeviceOnce attr_set;
    if (!attr_set.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(rowgemm_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(rowgemm_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set.set(dev);
    }
    const int blocks = (int)std::min<long long>(tiles, sms);
    if (p.Np > 128) rowgemm_ws_kernel<true><<<blocks, kRgThreads, smem, st>>>(p);
    else rowgemm_ws_kernel<false><<<blocks, kRgThreads, smem, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace gg

extern "C" int gridgcn_rowmlp_fwd(const float *in1, int ld1, int c1, const float *in2, int ld2, int c2,
                                  const float *weight, const float *bias, int cout, int relu_in,
                                  int relu_out, const float *row_scale, float *out, int ld_out,
                                  const float *cent, float *out_table, long long rows, void *stream);

extern "C" int gridgcn_rowmlp_tc_fwd(const float *in1, int ld1, int c1, const float *in2, int ld2, int c2,
                                     const float *weight, const float *bias, int cout, int relu_in,
                                     int relu_out, const float *row_scale, float *out, int ld_out,
                                     const float *cent, float *out_table, long long rows, void *stream) {
    if (!in1 || !weight || !bias || !out || c1 < 1 || c2 < 0 || cout < 1 || rows < 0) return GRIDGCN_EINVAL;
    if (c2 > 0 && !in2) return GRIDGCN_EINVAL;
    if (ld1 < c1 || (c2 > 0 && ld2 < c2) || ld_out < cout) return GRIDGCN_EINVAL;
    if (cent && (!out_table || (reinterpret_cast<uintptr_t>(cent) & 15) ||
                 (reinterpret_cast<uintptr_t>(out_table) & 15) || (ld_out & 3)))
        return GRIDGCN_EINVAL;
    const int rc = gg::launch_rowgemm_tc(in1, ld1, c1, in2, ld2, c2, weight, bias, cout, relu_in, relu_out, row_scale, out,
                                         ld_out, cent, out_table, rows, static_cast<cudaStream_t>(stream));
    if (rc >= 0) return rc;
    // shapes the tensor-core kernel does not take (unaligned views, widths not a multiple of 4): same operator on the CUDA cores
    return gridgcn_rowmlp_fwd(in1, ld1, c1, in2, ld2, c2, weight, bias, cout, relu_in, relu_out, row_scale, out, ld_out,
                              cent, out_table, rows, stream);
}
