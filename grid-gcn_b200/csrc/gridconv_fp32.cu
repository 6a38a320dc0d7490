// Fused GridConv layer, CUDA-core fp32 path (GRIDGCN_PRECISION_FP32) + the C-ABI dispatcher.
//
// A persistent CTA walks tiles of 64 edges (= 64/K centres; centres with K > 64 are processed in
// 64-row chunks with a running max).  Per tile: gather the neighbour rows of the previous layer's
// table with 128-bit loads straight into shared memory, build geo / attention inputs in registers,
// run the MLP chain with 4x4 register micro-tiles against weight chunks staged (transposed)
// through shared memory, multiply attention and feature activations in registers, max-pool over
// the K slots in shared memory and write one [cent | feats] row per centre.
#include "gridconv_common.cuh"

#include <algorithm>

namespace gg {

constexpr int kConvThreads = 256;
constexpr int kTileM = 64;
constexpr int kColBlk = 64;
constexpr int kKChunk = 32;
constexpr int kWsLd = kColBlk + 4;
constexpr int kPairLd = kColBlk + 1;

constexpr size_t kFp32SmemLimit = 220 * 1024;
constexpr int kFp32ScratchBlocks = 296;  // CTAs the global activation scratch is sized for (2 per SM)

// Activation buffers of one 64-edge tile (all [64 x ld] fp32):
//   XA / XB  feature MLP ping-pong (input row [geo | gathered features], hidden stages)
//   H0       [attention stage 0 | concat part]: with att_full the later attention stages read
//            [att_0 | feature MLP output ("next") or input ("last")], laid out side by side here
//   H1 / H2  ping-pong of the middle attention stages (classification block: 3 attention stages)
// They live in shared memory when everything fits (the segmentation block always does), else in a
// per-CTA slice of the global scratch (wide classification layers); ATT, the weight chunk, the pair
// tile and the running max stay in shared memory either way.
struct ConvSmemLayout {
    int xa, xb, h0, h1, h2;                     // offsets (floats) inside the activation region
    int attin, ws, pair, maxbuf, total;         // offsets (floats) inside shared memory; total = smem floats
    int ldx, ldh0, ldh1, cpt, Kc, nchunk;
    int a0w, extra;                             // width of attention stage 0, of the concat part
    int act_floats, act_global;
};

__host__ __device__ inline int pad4p(int w) { return ((w + 3) & ~3) + 4; }

__host__ __device__ inline ConvSmemLayout conv_smem_layout(const ConvParams &p) {
    ConvSmemLayout s;
    int wmax = p.feat_in;
    for (int i = 0; i + 1 < p.n_feat; i++) wmax = max(wmax, p.cout[i]);
    s.ldx = pad4p(wmax);
    s.a0w = p.n_att > 0 ? p.cout[p.n_feat] : 0;
    s.extra = p.att_full == GRIDGCN_ATT_FULL_NEXT ? p.Cout : (p.att_full == GRIDGCN_ATT_FULL_LAST ? p.feat_in : 0);
    s.ldh0 = p.n_att > 0 ? pad4p(s.a0w + s.extra) : 4;
    int mid = 0;
    for (int i = p.n_feat + 1; i + 1 < p.n_stages; i++) mid = max(mid, p.cout[i]);
    s.ldh1 = mid > 0 ? pad4p(mid) : 0;
    s.Kc = p.K < kTileM ? p.K : kTileM;
    s.nchunk = (p.K + kTileM - 1) / kTileM;
    s.cpt = kTileM / s.Kc;
    int off = 0;
    s.xa = off;  off += kTileM * s.ldx;
    s.xb = off;  off += (p.n_feat > 1 ? kTileM * s.ldx : 0);
    s.h0 = off;  off += kTileM * s.ldh0;
    s.h1 = off;  off += kTileM * s.ldh1;
    s.h2 = off;  off += (p.n_att > 3 ? kTileM * s.ldh1 : 0);
    s.act_floats = off;
    const int fixed = kTileM * 12 + kKChunk * kWsLd + kTileM * kPairLd + s.cpt * p.Cout;
    s.act_global = ((size_t)(s.act_floats + fixed) * sizeof(float) > kFp32SmemLimit) ? 1 : 0;
    off = s.act_global ? 0 : s.act_floats;
    s.attin = off;  off += kTileM * 12;
    s.ws = off;     off += kKChunk * kWsLd;
    s.pair = off;   off += kTileM * kPairLd;
    s.maxbuf = off; off += s.cpt * p.Cout;
    s.total = off;
    return s;
}

// acc[i][j] += sum_k in[(ty*4+i)*ld + k] * W[(cb*64 + tx*4 + j)*Cin + k]
__device__ __forceinline__ void dense_block(float acc[4][4], const float *in, int ld,
                                            const float *__restrict__ W, int Cin, int Cout, int cb,
                                            float *Ws, int tx, int ty) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < Cin; k0 += kKChunk) {
        __syncthreads();
        for (int e = tid; e < kColBlk * kKChunk; e += kConvThreads) {
            int c = e / kKChunk, kk = e % kKChunk;
            int col = cb * kColBlk + c, k = k0 + kk;
            Ws[kk * kWsLd + c] = (col < Cout && k < Cin) ? __ldg(W + (size_t)col * Cin + k) : 0.f;
        }
        __syncthreads();
        const int kend = min(kKChunk, Cin - k0);
        for (int kk = 0; kk < kend; kk++) {
            float4 w = *reinterpret_cast<const float4 *>(Ws + kk * kWsLd + tx * 4);
            float a[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = in[(ty * 4 + i) * ld + k0 + kk];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                acc[i][0] = fmaf(a[i], w.x, acc[i][0]);
                acc[i][1] = fmaf(a[i], w.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], w.z, acc[i][2]);
                acc[i][3] = fmaf(a[i], w.w, acc[i][3]);
            }
        }
    }
}

// out[r][col] = relu(acc + bias) for one hidden stage
__device__ __forceinline__ void hidden_stage(const float *in, int ld_in, float *out, int ld_out,
                                             const float *W, const float *bias, int Cin, int Cout,
                                             float *Ws, int tx, int ty) {
    for (int cb = 0; cb * kColBlk < Cout; cb++) {
        float acc[4][4];
        dense_block(acc, in, ld_in, W, Cin, Cout, cb, Ws, tx, ty);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int col = cb * kColBlk + tx * 4 + j;
            if (col < Cout) {
                float bj = __ldg(bias + col);
#pragma unroll
                for (int i = 0; i < 4; i++) out[(ty * 4 + i) * ld_out + col] = fmaxf(acc[i][j] + bj, 0.f);
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kConvThreads)
gridconv_fp32_kernel(ConvParams p, int num_tiles) {
    extern __shared__ __align__(16) float smem_f[];
    const ConvSmemLayout s = conv_smem_layout(p);
    float *act = s.act_global ? p.scratch + (size_t)blockIdx.x * s.act_floats : smem_f;
    float *XA = act + s.xa, *XB = act + s.xb, *H0 = act + s.h0, *H1 = act + s.h1, *H2 = act + s.h2;
    float *ATT = smem_f + s.attin;
    float *Ws = smem_f + s.ws, *PAIR = smem_f + s.pair, *MAXB = smem_f + s.maxbuf;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int row_w = 4 + p.Cin;
    const long long rows_total = (long long)p.B * p.Nprev;
    const long long centers_total = (long long)p.B * p.O;
    const int att_w = p.attfdim <= 0 ? 0 : (p.attfdim <= 3 ? 3 : (p.attfdim < 10 ? 4 : 10));
    const int gpre = p.Cin > 0 ? p.localfdim : 0;  // geo prefix in front of the gathered features
    const bool next = p.att_full == GRIDGCN_ATT_FULL_NEXT;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const long long c_base = (long long)tile * s.cpt;
        for (int i = tid; i < s.cpt * p.Cout; i += kConvThreads) MAXB[i] = -3.402823466e+38f;
        for (int chunk = 0; chunk < s.nchunk; chunk++) {
            __syncthreads();
            // ---- gather: one warp per edge row ----
            for (int r = warp; r < kTileM; r += kConvThreads / 32) {
                const int cl = r / s.Kc, pslot = chunk * kTileM + r % s.Kc;
                const long long center = c_base + cl;
                const bool valid = cl < s.cpt && center < centers_total && pslot < p.K;
                float *xrow = XA + r * s.ldx;
                if (!valid) {
                    for (int c = lane; c < s.ldx; c += 32) xrow[c] = 0.f;
                    if (lane < 12) ATT[r * 12 + lane] = 0.f;
                    continue;
                }
                const int b = (int)(center / p.O);
                const int idx = __ldg(p.nebidx + center * p.K + pslot);
                const float *src = p.table + take_row(idx, b, p.Nprev, rows_total) * row_w;
                float4 head;
                if ((row_w & 3) == 0) {
                    head = __ldg(reinterpret_cast<const float4 *>(src));
                } else {
                    head = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 0.f);
                }
                const float4 c4 = __ldg(p.cent + center);
                float att[10], dx, dy, dz;
                att_vector(p.attfdim, c4, head.x, head.y, head.z, att, dx, dy, dz);
                if (lane < att_w) {
                    float v = 0.f;
#pragma unroll
                    for (int q = 0; q < 10; q++) v = (lane == q) ? att[q] : v;
                    ATT[r * 12 + lane] = v;
                }
                if (p.Cin == 0 || gpre > 0) {  // geo vector: the whole input (has_feats == False, :242-243)
                    if (lane == 0) { xrow[0] = dx; xrow[1] = dy; xrow[2] = dz; }  // or its prefix (localfdim)
                }
                if (p.Cin > 0) {
                    if (gpre == 0 && (row_w & 3) == 0) {
                        const float4 *s4 = reinterpret_cast<const float4 *>(src) + 1;
                        for (int c = lane; c < p.Cin / 4; c += 32)
                            *reinterpret_cast<float4 *>(xrow + c * 4) = __ldg(s4 + c);
                    } else {
                        for (int c = lane; c < p.Cin; c += 32) xrow[gpre + c] = __ldg(src + 4 + c);
                    }
                }
            }
            __syncthreads();
            if (p.att_full == GRIDGCN_ATT_FULL_LAST) {  // concat part = the feature MLP's input
                for (int e = tid; e < kTileM * p.feat_in; e += kConvThreads) {
                    const int r = e / p.feat_in, c = e - r * p.feat_in;
                    H0[r * s.ldh0 + s.a0w + c] = XA[r * s.ldx + c];
                }
            }
            // ---- feature MLP hidden stages (with att_full "next" the last stage is materialised too,
            //      next to attention stage 0) ----
            const float *fin = XA;
            float *fout = XB;
            for (int st = 0; st + 1 < p.n_feat; st++) {
                hidden_stage(fin, s.ldx, fout, s.ldx, p.w[st], p.bias[st], p.cin[st], p.cout[st], Ws,
                             tx, ty);
                const float *t = fin;
                fin = fout;
                fout = const_cast<float *>(t);
            }
            const int sf = p.n_feat - 1, a0 = p.n_feat, sa = p.n_stages - 1;
            if (next)
                hidden_stage(fin, s.ldx, H0 + s.a0w, s.ldh0, p.w[sf], p.bias[sf], p.cin[sf], p.Cout, Ws, tx, ty);
            // ---- attention stages but the last ----
            const float *ain = H0;
            int ald = s.ldh0;
            if (p.n_att > 0) {
                hidden_stage(ATT, 12, H0, s.ldh0, p.w[a0], p.bias[a0], p.cin[a0], p.cout[a0], Ws, tx, ty);
                float *aout = H1;
                for (int st = a0 + 1; st < sa; st++) {
                    hidden_stage(ain, ald, aout, s.ldh1, p.w[st], p.bias[st], p.cin[st], p.cout[st], Ws, tx, ty);
                    ain = aout;
                    ald = s.ldh1;
                    aout = aout == H1 ? H2 : H1;
                }
            }
            // ---- last feature stage x last attention stage, product, max over K ----
            for (int cb = 0; cb * kColBlk < p.Cout; cb++) {
                float accf[4][4], acca[4][4];
                if (!next) dense_block(accf, fin, s.ldx, p.w[sf], p.cin[sf], p.Cout, cb, Ws, tx, ty);
                if (p.n_att > 0)
                    dense_block(acca, ain, ald, p.w[sa], p.cin[sa], p.Cout, cb, Ws, tx, ty);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int col = cb * kColBlk + tx * 4 + j;
                    float bf = (!next && col < p.Cout) ? __ldg(p.bias[sf] + col) : 0.f;
                    float ba = (p.n_att > 0 && col < p.Cout) ? __ldg(p.bias[sa] + col) : 0.f;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        float f;
                        if (next) f = col < p.Cout ? H0[(ty * 4 + i) * s.ldh0 + s.a0w + col] : 0.f;
                        else f = fmaxf(accf[i][j] + bf, 0.f);
                        if (p.n_att > 0) f *= fmaxf(acca[i][j] + ba, 0.f);  // :167 att * feats
                        PAIR[(ty * 4 + i) * kPairLd + tx * 4 + j] = f;
                    }
                }
                __syncthreads();
                for (int e = tid; e < s.cpt * kColBlk; e += kConvThreads) {
                    const int cl = e / kColBlk, c = e % kColBlk, col = cb * kColBlk + c;
                    if (col >= p.Cout || c_base + cl >= centers_total) continue;
                    const int nrow = min(s.Kc, p.K - chunk * kTileM);
                    float m = MAXB[cl * p.Cout + col];
                    for (int q = 0; q < nrow; q++) m = fmaxf(m, PAIR[(cl * s.Kc + q) * kPairLd + c]);
                    MAXB[cl * p.Cout + col] = m;
                }
                __syncthreads();
            }
        }
        // ---- epilogue: pre-ReLU, centre mask, [cent | feats] row ----
        const int out_w = 4 + p.Cout;
        for (int e = tid; e < s.cpt * out_w; e += kConvThreads) {
            const int cl = e / out_w, c = e % out_w;
            const long long center = c_base + cl;
            if (center >= centers_total) continue;
            float v;
            if (c < 4) {
                v = __ldg(reinterpret_cast<const float *>(p.cent + center) + c);
            } else {
                v = MAXB[cl * p.Cout + c - 4];
                if (p.pre_relu) v = fmaxf(v, 0.f);
                v *= __ldg(p.centmsk + center);
            }
            p.out[center * out_w + c] = v;
        }
        __syncthreads();
    }
}

// gridconv_tc.cu
int launch_gridconv_tc(const ConvParams &p, int precision, const float *packed, float *ftab, cudaStream_t st);
int tc_packed_floats(const ConvParams &c);
int tc_pack(const ConvParams &c, float *packed, cudaStream_t st);
void tc_set_phase_buffer(unsigned long long *buf);

static int launch_gridconv_fp32(const ConvParams &p, cudaStream_t st) {
    ConvSmemLayout s = conv_smem_layout(p);
    size_t smem = (size_t)s.total * sizeof(float);
    if (smem > kFp32SmemLimit) return GRIDGCN_ELIMIT;
    if (s.act_global && !p.scratch) return GRIDGCN_EWORKSPACE;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static PerDeviceOnce attr_set;
    if (!attr_set.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(gridconv_fp32_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFp32SmemLimit);
        if (e != cudaSuccess) return (int)e;
        attr_set.set(dev);
    }
    long long centers = (long long)p.B * p.O;
    long long tiles = (centers + s.cpt - 1) / s.cpt;
    if (tiles > 0x7fffffff) return GRIDGCN_ELIMIT;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = (int)max((size_t)1, min((size_t)8, kFp32SmemLimit / (smem + 1024)));
    int blocks = (int)min(tiles, (long long)sms * per_sm);
    if (s.act_global) blocks = min(blocks, kFp32ScratchBlocks);
    gridconv_fp32_kernel<<<blocks, kConvThreads, smem, st>>>(p, (int)tiles);
    return (int)cudaGetLastError();
}

// Validates an MLP description and fills the stage tables shared by every GridConv entry point.
static int fill_mlp(const gridgcn_mlp_t *m, int Cin, ConvParams &p) {
    if (!m || Cin < 0) return GRIDGCN_EINVAL;
    if (m->attfdim != 0 && m->attfdim != 3 && m->attfdim != 4 && m->attfdim != 10)
        return GRIDGCN_ELIMIT;
    if (m->localfdim != 0 && m->localfdim != 3) return GRIDGCN_ELIMIT;
    if (m->att_full < GRIDGCN_ATT_FULL_OFF || m->att_full > GRIDGCN_ATT_FULL_LAST) return GRIDGCN_EINVAL;
    p.Cin = Cin;
    p.n_feat = m->n_feat_stages;
    p.attfdim = m->attfdim;
    p.n_att = p.attfdim > 0 ? (m->n_att_stages > 0 ? m->n_att_stages : 2) : 0;
    if (p.n_feat < 1 || (p.attfdim > 0 && p.n_att < 2) || p.n_feat + p.n_att > GRIDGCN_MAX_STAGES)
        return GRIDGCN_EINVAL;
    p.localfdim = Cin > 0 ? m->localfdim : 0;  // without input features the MLP input is the geo vector anyway
    p.att_full = p.attfdim > 0 ? m->att_full : GRIDGCN_ATT_FULL_OFF;
    p.feat_in = Cin == 0 ? 3 : Cin + p.localfdim;
    if (m->feat_in != p.feat_in) return GRIDGCN_EINVAL;
    p.pre_relu = m->pre_relu;
    p.n_stages = p.n_feat + p.n_att;
    int win = p.feat_in;
    for (int i = 0; i < p.n_feat; i++) {
        p.cin[i] = win;
        p.cout[i] = m->widths[i];
        win = m->widths[i];
    }
    p.Cout = win;
    if (p.attfdim > 0) {
        const int a0 = p.n_feat;
        p.cin[a0] = att_in_width(p.attfdim);
        p.cout[a0] = m->widths[a0];
        win = m->widths[a0] + (p.att_full == GRIDGCN_ATT_FULL_NEXT ? p.Cout
                               : p.att_full == GRIDGCN_ATT_FULL_LAST ? p.feat_in : 0);
        for (int i = a0 + 1; i < p.n_stages; i++) {
            p.cin[i] = win;
            p.cout[i] = m->widths[i];
            win = m->widths[i];
        }
        if (win != p.Cout) return GRIDGCN_EINVAL;  // att * feats needs equal widths
    }
    for (int i = 0; i < p.n_stages; i++) {
        if (p.cout[i] < 1 || p.cout[i] > 1024 || !m->weight[i] || !m->bias[i]) return GRIDGCN_EINVAL;
        p.w[i] = m->weight[i];
        p.bias[i] = m->bias[i];
    }
    return 0;
}

// The tensor-core kernels implement the segmentation block: gathered features as they are, two attention
// stages, no concat.
static bool tc_supported(const ConvParams &p) {
    return p.localfdim == 0 && p.att_full == GRIDGCN_ATT_FULL_OFF && (p.attfdim == 0 || p.n_att == 2);
}

// gridconv_cls_tc.cu: the classification-block variants as a chain of tensor-core row GEMMs over the edge rows
bool cls_tc_ok(const ConvParams &p);
long long cls_tc_packed_floats(const ConvParams &p);
long long cls_tc_edge_bytes(const ConvParams &p);
int cls_tc_pack(const ConvParams &p, float *packed, cudaStream_t st);
int launch_gridconv_cls_tc(const ConvParams &p, const float *packed, float *ws, size_t ws_bytes, cudaStream_t st);
constexpr size_t kClsWorkspaceCap = (size_t)6 << 30;  // clouds are processed in chunks that fit this much workspace

}  // namespace gg

using namespace gg;

extern "C" size_t gridgcn_gridconv_packed_bytes(const gridgcn_mlp_t *m, int Cin) {
    ConvParams p{};
    if (fill_mlp(m, Cin, p)) return 0;
    if (!tc_supported(p)) return cls_tc_ok(p) ? (size_t)cls_tc_packed_floats(p) * sizeof(float) : 0;
    int n = tc_packed_floats(p);
    return n < 0 ? 0 : (size_t)n * sizeof(float);
}

extern "C" int gridgcn_gridconv_pack(const gridgcn_mlp_t *m, int Cin, void *packed, size_t packed_bytes,
                                     void *stream) {
    ConvParams p{};
    int rc = fill_mlp(m, Cin, p);
    if (rc) return rc;
    if (!tc_supported(p)) {
        if (!cls_tc_ok(p)) return GRIDGCN_ELIMIT;
        if (!packed || packed_bytes < (size_t)cls_tc_packed_floats(p) * sizeof(float) ||
            (reinterpret_cast<uintptr_t>(packed) & 15))
            return GRIDGCN_EWORKSPACE;
        return cls_tc_pack(p, static_cast<float *>(packed), static_cast<cudaStream_t>(stream));
    }
    int n = tc_packed_floats(p);
    if (n < 0) return GRIDGCN_ELIMIT;
    if (!packed || packed_bytes < (size_t)n * sizeof(float) || (reinterpret_cast<uintptr_t>(packed) & 15))
        return GRIDGCN_EWORKSPACE;
    return tc_pack(p, static_cast<float *>(packed), static_cast<cudaStream_t>(stream));
}

extern "C" size_t gridgcn_gridconv_workspace_bytes(const gridgcn_mlp_t *m, int B, int Nprev, int Cin) {
    ConvParams p{};
    if (fill_mlp(m, Cin, p) || B < 0 || Nprev < 0 || !tc_supported(p)) return 0;
    if (Cin <= 0) return 0;
    // the transformed feature table F (rows x Cout) + two ping-pong buffers for the hidden activations of the
    // per-point feature MLP when it runs as a chain of row GEMMs (gridconv_tc.cu, rowgemm_tc.cu)
    int hmax = 0;
    for (int s = 0; s + 1 < p.n_feat; s++) hmax = p.cout[s] > hmax ? p.cout[s] : hmax;
    return (size_t)B * Nprev * ((size_t)p.Cout + 2 * (size_t)hmax) * sizeof(float);
}

extern "C" size_t gridgcn_gridconv_edge_workspace_bytes(const gridgcn_mlp_t *m, int B, int Cin, int O, int K) {
    ConvParams p{};
    if (fill_mlp(m, Cin, p) || B < 1 || O < 1 || K < 1 || tc_supported(p) || !cls_tc_ok(p)) return 0;
    const size_t per_cloud = (size_t)O * K * (size_t)cls_tc_edge_bytes(p);
    const size_t clouds = std::max<size_t>(1, std::min<size_t>((size_t)B, kClsWorkspaceCap / per_cloud));
    return clouds * per_cloud;
}

extern "C" size_t gridgcn_gridconv_fp32_scratch_bytes(const gridgcn_mlp_t *m, int Cin, int K) {
    ConvParams p{};
    if (fill_mlp(m, Cin, p) || K < 1) return 0;
    p.K = K;
    const ConvSmemLayout s = conv_smem_layout(p);
    return s.act_global ? (size_t)kFp32ScratchBlocks * s.act_floats * sizeof(float) : 0;
}

extern "C" int gridgcn_gridconv_fwd(const float *table, const int *nebidx, const float *cent,
                                    const float *centmsk, int B, int Nprev, int Cin, int O, int K,
                                    const gridgcn_mlp_t *m, int precision, const void *packed,
                                    void *workspace, size_t workspace_bytes, float *out,
                                    void *stream) {
    if (!table || !nebidx || !cent || !centmsk || !out || !m) return GRIDGCN_EINVAL;
    if (B < 0 || Nprev < 1 || Cin < 0 || O < 1 || K < 1) return GRIDGCN_EINVAL;
    if (K > 1024) return GRIDGCN_ELIMIT;
    if ((reinterpret_cast<uintptr_t>(table) & 15) || (reinterpret_cast<uintptr_t>(cent) & 15))
        return GRIDGCN_EINVAL;
    ConvParams p{};
    int rc = fill_mlp(m, Cin, p);
    if (rc) return rc;
    p.table = table;
    p.nebidx = nebidx;
    p.cent = reinterpret_cast<const float4 *>(cent);
    p.centmsk = centmsk;
    p.out = out;
    p.B = B; p.Nprev = Nprev; p.O = O; p.K = K;
    if (B == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (precision == GRIDGCN_PRECISION_FP32) {
        const size_t need = gridgcn_gridconv_fp32_scratch_bytes(m, Cin, K);
        if (need && (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15)))
            return GRIDGCN_EWORKSPACE;
        p.scratch = need ? static_cast<float *>(workspace) : nullptr;
        return launch_gridconv_fp32(p, st);
    }
    if (precision == GRIDGCN_PRECISION_TF32 || precision == GRIDGCN_PRECISION_TF32X3) {
        if (!packed || (reinterpret_cast<uintptr_t>(packed) & 15)) return GRIDGCN_EWORKSPACE;
        if (!tc_supported(p)) {  // classification-block variants: un-fused chain of tensor-core row GEMMs (3-pass only)
            if (precision != GRIDGCN_PRECISION_TF32X3 || !cls_tc_ok(p)) return GRIDGCN_ELIMIT;
            if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 15)) return GRIDGCN_EWORKSPACE;
            return launch_gridconv_cls_tc(p, static_cast<const float *>(packed), static_cast<float *>(workspace),
                                          workspace_bytes, st);
        }
        if (workspace_bytes < gridgcn_gridconv_workspace_bytes(m, B, Nprev, Cin) ||
            (Cin > 0 && (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 15))))
            return GRIDGCN_EWORKSPACE;
        return launch_gridconv_tc(p, precision, static_cast<const float *>(packed),
                                  static_cast<float *>(workspace), st);
    }
    return GRIDGCN_EINVAL;
}

// Debug: 8 x u64 device buffer that CTA 0 of the per-edge tensor-core kernel fills with the cycles it
// spent in each phase (gather, sync, hidden MMA, hidden epilogue, sync, final MMA, final epilogue,
// sync).  Pass NULL to switch the accounting off.  Not re-entrant; not part of the operator ABI.
extern "C" void gridgcn_debug_phase_buffer(unsigned long long *dev_buf) { tc_set_phase_buffer(dev_buf); }
