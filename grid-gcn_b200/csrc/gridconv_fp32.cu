// Fused GridConv layer, CUDA-core fp32 path (GRIDGCN_PRECISION_FP32) + the C-ABI dispatcher.
//
// A persistent CTA walks tiles of 64 edges (= 64/K centres; centres with K > 64 are processed in
// 64-row chunks with a running max).  Per tile: gather the neighbour rows of the previous layer's
// table with 128-bit loads straight into shared memory, build geo / attention inputs in registers,
// run the MLP chain with 4x4 register micro-tiles against weight chunks staged (transposed)
// through shared memory, multiply attention and feature activations in registers, max-pool over
// the K slots in shared memory and write one [cent | feats] row per centre.
#include "gridconv_common.cuh"

namespace gg {

constexpr int kConvThreads = 256;
constexpr int kTileM = 64;
constexpr int kColBlk = 64;
constexpr int kKChunk = 32;
constexpr int kWsLd = kColBlk + 4;
constexpr int kPairLd = kColBlk + 1;

struct ConvSmemLayout {
    int x0, x1, h, attin, ws, pair, maxbuf, total;  // offsets in floats
    int ldx, ldh, cpt, Kc, nchunk;
};

__host__ __device__ inline ConvSmemLayout conv_smem_layout(const ConvParams &p) {
    ConvSmemLayout s;
    int wmax = p.feat_in;
    for (int i = 0; i + 1 < p.n_feat; i++) wmax = max(wmax, p.cout[i]);
    s.ldx = ((wmax + 3) & ~3) + 4;
    s.ldh = p.attfdim > 0 ? ((p.cout[p.n_feat] + 3) & ~3) + 4 : 4;
    s.Kc = p.K < kTileM ? p.K : kTileM;
    s.nchunk = (p.K + kTileM - 1) / kTileM;
    s.cpt = kTileM / s.Kc;
    int off = 0;
    s.x0 = off;     off += kTileM * s.ldx;
    s.x1 = off;     off += (p.n_feat > 1 ? kTileM * s.ldx : 0);
    s.h = off;      off += kTileM * s.ldh;
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
    float *X0 = smem_f + s.x0, *X1 = smem_f + s.x1, *H = smem_f + s.h, *ATT = smem_f + s.attin;
    float *Ws = smem_f + s.ws, *PAIR = smem_f + s.pair, *MAXB = smem_f + s.maxbuf;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int row_w = 4 + p.Cin;
    const long long rows_total = (long long)p.B * p.Nprev;
    const long long centers_total = (long long)p.B * p.O;
    const int att_w = p.attfdim <= 0 ? 0 : (p.attfdim <= 3 ? 3 : (p.attfdim < 10 ? 4 : 10));

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
                float *xrow = X0 + r * s.ldx;
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
                if (p.Cin == 0) {  // has_feats == False: features are the geo vector (:242-243)
                    if (lane == 0) { xrow[0] = dx; xrow[1] = dy; xrow[2] = dz; }
                } else if ((row_w & 3) == 0) {
                    const float4 *s4 = reinterpret_cast<const float4 *>(src) + 1;
                    for (int c = lane; c < p.Cin / 4; c += 32)
                        *reinterpret_cast<float4 *>(xrow + c * 4) = __ldg(s4 + c);
                } else {
                    for (int c = lane; c < p.Cin; c += 32) xrow[c] = __ldg(src + 4 + c);
                }
            }
            __syncthreads();
            // ---- feature MLP hidden stages ----
            const float *fin = X0;
            float *fout = X1;
            for (int st = 0; st + 1 < p.n_feat; st++) {
                hidden_stage(fin, s.ldx, fout, s.ldx, p.w[st], p.bias[st], p.cin[st], p.cout[st], Ws,
                             tx, ty);
                const float *t = fin;
                fin = fout;
                fout = const_cast<float *>(t);
            }
            // ---- attention hidden stage ----
            if (p.attfdim > 0)
                hidden_stage(ATT, 12, H, s.ldh, p.w[p.n_feat], p.bias[p.n_feat], p.cin[p.n_feat],
                             p.cout[p.n_feat], Ws, tx, ty);
            // ---- last feature stage x last attention stage, product, max over K ----
            const int sf = p.n_feat - 1, sa = p.n_feat + 1;
            for (int cb = 0; cb * kColBlk < p.Cout; cb++) {
                float accf[4][4], acca[4][4];
                dense_block(accf, fin, s.ldx, p.w[sf], p.cin[sf], p.Cout, cb, Ws, tx, ty);
                if (p.attfdim > 0)
                    dense_block(acca, H, s.ldh, p.w[sa], p.cin[sa], p.Cout, cb, Ws, tx, ty);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int col = cb * kColBlk + tx * 4 + j;
                    float bf = col < p.Cout ? __ldg(p.bias[sf] + col) : 0.f;
                    float ba = (p.attfdim > 0 && col < p.Cout) ? __ldg(p.bias[sa] + col) : 0.f;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        float f = fmaxf(accf[i][j] + bf, 0.f);
                        if (p.attfdim > 0) f *= fmaxf(acca[i][j] + ba, 0.f);  // :167 att * feats
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
    if (smem > 220 * 1024) return GRIDGCN_ELIMIT;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gridconv_fp32_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    long long centers = (long long)p.B * p.O;
    long long tiles = (centers + s.cpt - 1) / s.cpt;
    if (tiles > 0x7fffffff) return GRIDGCN_ELIMIT;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = (int)max((size_t)1, min((size_t)8, (size_t)(220 * 1024) / (smem + 1024)));
    int blocks = (int)min(tiles, (long long)sms * per_sm);
    gridconv_fp32_kernel<<<blocks, kConvThreads, smem, st>>>(p, (int)tiles);
    return (int)cudaGetLastError();
}

// Validates an MLP description and fills the stage tables shared by every GridConv entry point.
static int fill_mlp(const gridgcn_mlp_t *m, int Cin, ConvParams &p) {
    if (!m || Cin < 0) return GRIDGCN_EINVAL;
    if (m->n_feat_stages < 1 || m->n_feat_stages > GRIDGCN_MAX_STAGES - 2) return GRIDGCN_EINVAL;
    if (m->attfdim != 0 && m->attfdim != 3 && m->attfdim != 4 && m->attfdim != 10)
        return GRIDGCN_ELIMIT;
    p.Cin = Cin;
    p.n_feat = m->n_feat_stages;
    p.attfdim = m->attfdim;
    p.feat_in = Cin == 0 ? 3 : Cin;
    if (m->feat_in != p.feat_in) return GRIDGCN_EINVAL;
    p.pre_relu = m->pre_relu;
    p.n_stages = p.n_feat + (p.attfdim > 0 ? 2 : 0);
    int win = p.feat_in;
    for (int i = 0; i < p.n_feat; i++) {
        p.cin[i] = win;
        p.cout[i] = m->widths[i];
        win = m->widths[i];
    }
    p.Cout = win;
    if (p.attfdim > 0) {
        p.cin[p.n_feat] = att_in_width(p.attfdim);
        p.cout[p.n_feat] = m->widths[p.n_feat];
        p.cin[p.n_feat + 1] = m->widths[p.n_feat];
        p.cout[p.n_feat + 1] = m->widths[p.n_feat + 1];
        if (p.cout[p.n_feat + 1] != p.Cout) return GRIDGCN_EINVAL;  // att * feats needs equal widths
    }
    for (int i = 0; i < p.n_stages; i++) {
        if (p.cout[i] < 1 || p.cout[i] > 1024 || !m->weight[i] || !m->bias[i]) return GRIDGCN_EINVAL;
        p.w[i] = m->weight[i];
        p.bias[i] = m->bias[i];
    }
    return 0;
}

}  // namespace gg

using namespace gg;

extern "C" size_t gridgcn_gridconv_packed_bytes(const gridgcn_mlp_t *m, int Cin) {
    ConvParams p{};
    if (fill_mlp(m, Cin, p)) return 0;
    int n = tc_packed_floats(p);
    return n < 0 ? 0 : (size_t)n * sizeof(float);
}

extern "C" int gridgcn_gridconv_pack(const gridgcn_mlp_t *m, int Cin, void *packed, size_t packed_bytes,
                                     void *stream) {
    ConvParams p{};
    int rc = fill_mlp(m, Cin, p);
    if (rc) return rc;
    int n = tc_packed_floats(p);
    if (n < 0) return GRIDGCN_ELIMIT;
    if (!packed || packed_bytes < (size_t)n * sizeof(float) || (reinterpret_cast<uintptr_t>(packed) & 15))
        return GRIDGCN_EWORKSPACE;
    return tc_pack(p, static_cast<float *>(packed), static_cast<cudaStream_t>(stream));
}

extern "C" size_t gridgcn_gridconv_workspace_bytes(const gridgcn_mlp_t *m, int B, int Nprev, int Cin) {
    ConvParams p{};
    if (fill_mlp(m, Cin, p) || B < 0 || Nprev < 0) return 0;
    return Cin > 0 ? (size_t)B * Nprev * p.Cout * sizeof(float) : 0;
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
    if (precision == GRIDGCN_PRECISION_FP32) return launch_gridconv_fp32(p, st);
    if (precision == GRIDGCN_PRECISION_TF32 || precision == GRIDGCN_PRECISION_TF32X3) {
        if (!packed || (reinterpret_cast<uintptr_t>(packed) & 15)) return GRIDGCN_EWORKSPACE;
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
