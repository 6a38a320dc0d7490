// Training-mode GridConv block: the hand-written kernels of the forward (BatchNorm with BATCH statistics between the
// 1x1 convolutions) and of the backward pass.
//
// Reference: utils/ops.py:149-158 (conv2d: Convolution -> BatchNorm(fix_gamma=False, use_global_stats=False,
// momentum=bn_decay) -> relu; `use_global_stats: False` in segmentation/configs/configs.yaml:27),
// segmentation/models/gcn_module_g_att.py:120-170 (feature MLP, attention MLP, product), :45-79 (max pool over the K
// slots), :24-43 (pre-ReLU), :284-285 (centre mask), utils/ops.py:78-93 (batch_take_g), base_solver.py:153-156 (fit).
// The index operators in front (Gridify / GridifyKNN / BallKNN) have no gradient in the reference either
// (gridify-inl.h:227-231), so coordinates are constants and the gradient flows through the feature columns only.
//
// Layout: the block runs on EDGE ROWS (one row per (centre, slot) edge, channels contiguous) -- the layout of the
// tensor-core row GEMM (rowgemm_tc.cu), which computes every 1x1 convolution of the forward (Z = X W^T + b) and every
// input gradient of the backward (dX = dZ W); the kernels here are everything around those GEMMs:
//     edge rows + gather indices, per-channel batch statistics, normalise + ReLU, max pool with arg-max, its routing
//     backward, BatchNorm + ReLU backward (two passes: ReLU backward + per-channel sums, then elementwise), weight and
//     bias gradient (dW = dZ^T X, register-tiled), and the scatter-add of the gather.
// They are orchestrated from the host (grid-gcn_b200/train_cuda.py); this is an un-fused training path -- one HBM
// round trip per operator, like the reference's MXNet graph -- whose GEMMs run on tcgen05.
#include "gridconv_common.cuh"

namespace gg {

constexpr int kTrThreads = 256;

// per edge: neighbour row index (take with clip after the batch offset), X_f = [geo, 0] (Cin == 0) or the gathered
// feature columns, X_a = att_vec (padded to a multiple of 4)
__global__ void __launch_bounds__(kTrThreads)
train_edge_rows_kernel(const float *__restrict__ table, const int *__restrict__ nebidx, const float4 *__restrict__ cent,
                       int B, int Nprev, int Cin, int O, int K, int attfdim, int ain_p, float *__restrict__ xf,
                       float *__restrict__ xa, int *__restrict__ rowidx) {
    const int fin_p = Cin > 0 ? Cin : 4, groups = fin_p / 4, row_w = 4 + Cin;
    const long long edges = (long long)B * O * K, rows_total = (long long)B * Nprev;
    for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < edges * groups; i += (long long)gridDim.x * kTrThreads) {
        const long long e = i / groups;
        const int g = (int)(i % groups);
        const long long centre = e / K;
        const int b = (int)(centre / O);
        const long long row = take_row(__ldg(nebidx + e), b, Nprev, rows_total);
        const float *src = table + row * row_w;
        if (g == 0) {
            const float4 c = __ldg(cent + centre);
            float att[12];
#pragma unroll
            for (int k = 0; k < 12; k++) att[k] = 0.f;
            float dx, dy, dz;
            att_vector(attfdim > 0 ? attfdim : 3, c, __ldg(src), __ldg(src + 1), __ldg(src + 2), att, dx, dy, dz);
            for (int k = 0; k < ain_p; k += 4)
                *reinterpret_cast<float4 *>(xa + e * ain_p + k) = make_float4(att[k], att[k + 1], att[k + 2], att[k + 3]);
            rowidx[e] = (int)row;
            if (Cin == 0) {
                *reinterpret_cast<float4 *>(xf + e * 4) = make_float4(dx, dy, dz, 0.f);
                continue;
            }
        }
        *reinterpret_cast<float4 *>(xf + e * fin_p + 4 * g) = __ldg(reinterpret_cast<const float4 *>(src + 4 + 4 * g));
    }
}

// per-channel sums over the rows: s0[c] += sum a[r,c] * (b ? b[r,c] : 1), s1[c] += sum a[r,c]^2 (s1 may be null).
// One CTA per slab of rows, thread = channel (coalesced), one atomicAdd per channel and CTA.
__global__ void __launch_bounds__(kTrThreads)
train_col_sums_kernel(const float *__restrict__ a, const float *__restrict__ b, long long rows, int C, int rows_per_cta,
                      float *__restrict__ s0, float *__restrict__ s1) {
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    for (int c = threadIdx.x; c < C; c += kTrThreads) {
        float t0 = 0.f, t1 = 0.f;
        for (long long r = r0; r < r1; r++) {
            const float v = a[r * C + c];
            t0 += b ? v * b[r * C + c] : v;
            t1 += v * v;
        }
        atomicAdd(s0 + c, t0);
        if (s1) atomicAdd(s1 + c, t1);
    }
}

// batch statistics from the column sums: mean, biased variance -> invstd; moving statistics (torch convention:
// momentum weighs the NEW value, unbiased variance)
__global__ void __launch_bounds__(kTrThreads)
train_bn_finalize_kernel(const float *__restrict__ s0, const float *__restrict__ s1, float rows, int C, float eps, float momentum,
                         float *__restrict__ mean, float *__restrict__ invstd, float *__restrict__ running_mean,
                         float *__restrict__ running_var, long long *__restrict__ num_batches) {
    const int c = blockIdx.x * kTrThreads + threadIdx.x;
    if (c == 0 && num_batches) *num_batches += 1;
    if (c >= C) return;
    const float mu = s0[c] / rows;
    const float var = fmaxf(s1[c] / rows - mu * mu, 0.f);
    mean[c] = mu;
    invstd[c] = rsqrtf(var + eps);
    if (running_mean) running_mean[c] = running_mean[c] * (1.f - momentum) + momentum * mu;
    if (running_var) running_var[c] = running_var[c] * (1.f - momentum) + momentum * (var * (rows / fmaxf(rows - 1.f, 1.f)));
}

// y = relu(gamma * (z - mean) * invstd + beta)
__global__ void __launch_bounds__(kTrThreads)
train_bn_relu_fwd_kernel(const float *__restrict__ z, long long n, int C, const float *__restrict__ mean,
                         const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta,
                         float *__restrict__ y) {
    for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTrThreads) {
        const int c = (int)(i % C);
        y[i] = fmaxf(gamma[c] * (z[i] - mean[c]) * invstd[c] + beta[c], 0.f);
    }
}

// dz = dy * (y > 0);  xhat = (z - mean) * invstd;  writes dz in place of dy and xhat in place of z (scratch reuse)
__global__ void __launch_bounds__(kTrThreads)
train_relu_bwd_xhat_kernel(float *__restrict__ dy, const float *__restrict__ y, float *__restrict__ z, long long n, int C,
                           const float *__restrict__ mean, const float *__restrict__ invstd) {
    for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTrThreads) {
        const int c = (int)(i % C);
        dy[i] = y[i] > 0.f ? dy[i] : 0.f;
        z[i] = (z[i] - mean[c]) * invstd[c];
    }
}

// BatchNorm backward (batch statistics): dzpre = gamma * invstd / N * (N * dz - sum_dz - xhat * sum_dz_xhat)
__global__ void __launch_bounds__(kTrThreads)
train_bn_bwd_kernel(const float *__restrict__ dz, const float *__restrict__ xhat, long long n, int C, float inv_rows,
                    const float *__restrict__ gamma, const float *__restrict__ invstd, const float *__restrict__ sum_dz,
                    const float *__restrict__ sum_dzx, float *__restrict__ dzpre) {
    for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTrThreads) {
        const int c = (int)(i % C);
        dzpre[i] = gamma[c] * invstd[c] * (dz[i] - inv_rows * (sum_dz[c] + xhat[i] * sum_dzx[c]));
    }
}

// max pool over the K slots of F * A (A may be null), arg-max slot, pre-ReLU, centre mask
__global__ void __launch_bounds__(kTrThreads)
train_pool_fwd_kernel(const float *__restrict__ F, const float *__restrict__ A, long long centres, int K, int C, int pre_relu,
                      const float *__restrict__ mask, float *__restrict__ out, int ldo, int *__restrict__ argmax) {
    for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < centres * C; i += (long long)gridDim.x * kTrThreads) {
        const long long o = i / C;
        const int c = (int)(i % C);
        float m = -3.402823466e+38f;
        int am = 0;
        for (int k = 0; k < K; k++) {
            const long long e = (o * K + k) * C + c;
            const float v = A ? F[e] * A[e] : F[e];
            if (v > m) { m = v; am = k; }
        }
        if (pre_relu && m <= 0.f) { m = 0.f; am = -1; }  // relu gate closed: no gradient
        out[o * ldo + c] = m * mask[o];
        argmax[i] = am;
    }
}

// routing backward of the pool: dP goes to the arg-max edge only; dF = dP * A, dA = dP * F (zero elsewhere)
__global__ void __launch_bounds__(kTrThreads)
train_pool_bwd_kernel(const float *__restrict__ dout, int ldo, const float *__restrict__ F, const float *__restrict__ A,
                      const int *__restrict__ argmax, const float *__restrict__ mask, long long centres, int K, int C,
                      float *__restrict__ dF, float *__restrict__ dA) {
    for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < centres * K * C; i += (long long)gridDim.x * kTrThreads) {
        const int c = (int)(i % C);
        const long long ek = i / C, o = ek / K;
        const int k = (int)(ek % K);
        float g = 0.f;
        if (argmax[o * C + c] == k) g = dout[o * ldo + c] * mask[o];
        dF[i] = A ? g * A[i] : g;
        if (dA) dA[i] = g * F[i];
    }
}

// weight gradient dW[o, i] += sum_r dz[r, o] * x[r, i] over a slab of rows per CTA (x = [in1 | in2] row views).
// Tile: 32 x 32 outputs per CTA column/row block, rows streamed through shared memory.
__global__ void __launch_bounds__(kTrThreads)
train_wgrad_kernel(const float *__restrict__ dz, int Cout, const float *__restrict__ in1, int ld1, int c1,
                   const float *__restrict__ in2, int ld2, int c2, long long rows, int rows_per_cta, float *__restrict__ dW) {
    __shared__ float sd[32][33], sx[32][33];
    const int Kin = c1 + c2;
    const int o0 = blockIdx.y * 32, i0 = blockIdx.z * 32;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long rb = r0; rb < r1; rb += 32) {
        for (int j = ty; j < 32; j += 8) {
            const long long r = rb + j;
            const int o = o0 + tx, ii = i0 + tx;
            sd[j][tx] = (r < r1 && o < Cout) ? dz[r * Cout + o] : 0.f;
            float xv = 0.f;
            if (r < r1 && ii < Kin) xv = ii < c1 ? in1[r * ld1 + ii] : in2[r * ld2 + (ii - c1)];
            sx[j][tx] = xv;
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < 32; j++) {
            const float xv = sx[j][tx];
#pragma unroll
            for (int q = 0; q < 4; q++) acc[q] = fmaf(sd[j][ty + 8 * q], xv, acc[q]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int o = o0 + ty + 8 * q, ii = i0 + tx;
        if (o < Cout && ii < Kin) atomicAdd(dW + (size_t)o * Kin + ii, acc[q]);
    }
}


// ---- float4 forms (C % 4 == 0, C <= 1024): thread = one group of 4 consecutive channels, walking the rows of its CTA's
// slab with stride R = 256 / (C/4) -- 128-bit coalesced accesses, per-channel constants in registers, no div / mod per
// element.  The reductions end in one shared-memory pass and one atomicAdd per channel and CTA.
struct ColMap {
    int G, R, cg, rr;
    bool active;
    __device__ __forceinline__ ColMap(int C) {
        G = C >> 2;
        R = kTrThreads / G;
        cg = threadIdx.x % G;
        rr = threadIdx.x / G;
        active = rr < R;
    }
};

__device__ __forceinline__ void col_reduce_store(const ColMap &m, int C, float4 t, float (*red)[4], float *__restrict__ dst) {
    __syncthreads();
    red[threadIdx.x][0] = t.x; red[threadIdx.x][1] = t.y; red[threadIdx.x][2] = t.z; red[threadIdx.x][3] = t.w;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kTrThreads) {
        float acc = 0.f;
        for (int r = 0; r < m.R; r++) acc += red[r * m.G + (c >> 2)][c & 3];
        atomicAdd(dst + c, acc);
    }
}

template <bool HAS_B, bool SQ>
__global__ void __launch_bounds__(kTrThreads)
train_col_sums4_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b, long long rows, int C, int rows_per_cta,
                       float *__restrict__ s0, float *__restrict__ s1) {
    __shared__ float red[kTrThreads][4];
    const ColMap m(C);
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
    if (m.active)
        for (long long r = r0 + m.rr; r < r1; r += m.R) {
            const float4 v = __ldg(a + r * m.G + m.cg);
            if (HAS_B) {
                const float4 w = __ldg(b + r * m.G + m.cg);
                t0.x = fmaf(v.x, w.x, t0.x); t0.y = fmaf(v.y, w.y, t0.y); t0.z = fmaf(v.z, w.z, t0.z); t0.w = fmaf(v.w, w.w, t0.w);
            } else {
                t0.x += v.x; t0.y += v.y; t0.z += v.z; t0.w += v.w;
            }
            if (SQ) { t1.x = fmaf(v.x, v.x, t1.x); t1.y = fmaf(v.y, v.y, t1.y); t1.z = fmaf(v.z, v.z, t1.z); t1.w = fmaf(v.w, v.w, t1.w); }
        }
    col_reduce_store(m, C, t0, red, s0);
    if (SQ) col_reduce_store(m, C, t1, red, s1);
}

__global__ void __launch_bounds__(kTrThreads)
train_bn_relu_fwd4_kernel(const float4 *__restrict__ z, long long rows, int C, int rows_per_cta, const float4 *__restrict__ mean,
                          const float4 *__restrict__ invstd, const float4 *__restrict__ gamma, const float4 *__restrict__ beta,
                          float4 *__restrict__ y) {
    const ColMap m(C);
    if (!m.active) return;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    const float4 mu = __ldg(mean + m.cg), is = __ldg(invstd + m.cg), ga = __ldg(gamma + m.cg), be = __ldg(beta + m.cg);
    for (long long r = r0 + m.rr; r < r1; r += m.R) {
        const float4 v = __ldg(z + r * m.G + m.cg);
        float4 o;  // same association as the scalar kernel: gamma * (z - mean) * invstd + beta
        o.x = fmaxf(ga.x * (v.x - mu.x) * is.x + be.x, 0.f);
        o.y = fmaxf(ga.y * (v.y - mu.y) * is.y + be.y, 0.f);
        o.z = fmaxf(ga.z * (v.z - mu.z) * is.z + be.z, 0.f);
        o.w = fmaxf(ga.w * (v.w - mu.w) * is.w + be.w, 0.f);
        y[r * m.G + m.cg] = o;
    }
}

// ReLU backward + xhat (in place) AND the two BatchNorm-backward sums in the same pass:
//     dy <- dz = dy * (y > 0),  z <- xhat = (z - mean) * invstd,  sum_dz[c] += dz,  sum_dzx[c] += dz * xhat
__global__ void __launch_bounds__(kTrThreads)
train_relu_bwd_xhat4_kernel(float4 *__restrict__ dy, const float4 *__restrict__ y, float4 *__restrict__ z, long long rows, int C,
                            int rows_per_cta, const float4 *__restrict__ mean, const float4 *__restrict__ invstd,
                            float *__restrict__ sum_dz, float *__restrict__ sum_dzx) {
    __shared__ float red[kTrThreads][4];
    const ColMap m(C);
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
    if (m.active) {
        const float4 mu = __ldg(mean + m.cg), is = __ldg(invstd + m.cg);
        for (long long r = r0 + m.rr; r < r1; r += m.R) {
            const long long i = r * m.G + m.cg;
            float4 d = dy[i], x = z[i];
            const float4 yy = __ldg(y + i);
            d.x = yy.x > 0.f ? d.x : 0.f; d.y = yy.y > 0.f ? d.y : 0.f; d.z = yy.z > 0.f ? d.z : 0.f; d.w = yy.w > 0.f ? d.w : 0.f;
            x.x = (x.x - mu.x) * is.x; x.y = (x.y - mu.y) * is.y; x.z = (x.z - mu.z) * is.z; x.w = (x.w - mu.w) * is.w;
            dy[i] = d;
            z[i] = x;
            t0.x += d.x; t0.y += d.y; t0.z += d.z; t0.w += d.w;
            t1.x = fmaf(d.x, x.x, t1.x); t1.y = fmaf(d.y, x.y, t1.y); t1.z = fmaf(d.z, x.z, t1.z); t1.w = fmaf(d.w, x.w, t1.w);
        }
    }
    if (sum_dz) {
        col_reduce_store(m, C, t0, red, sum_dz);
        col_reduce_store(m, C, t1, red, sum_dzx);
    }
}

__global__ void __launch_bounds__(kTrThreads)
train_bn_bwd4_kernel(const float4 *__restrict__ dz, const float4 *__restrict__ xhat, long long rows, int C, int rows_per_cta,
                     float inv_rows, const float4 *__restrict__ gamma, const float4 *__restrict__ invstd,
                     const float4 *__restrict__ sum_dz, const float4 *__restrict__ sum_dzx, float4 *__restrict__ dzpre) {
    const ColMap m(C);
    if (!m.active) return;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    const float4 ga = __ldg(gamma + m.cg), is = __ldg(invstd + m.cg), sd = __ldg(sum_dz + m.cg), sx = __ldg(sum_dzx + m.cg);
    for (long long r = r0 + m.rr; r < r1; r += m.R) {
        const long long i = r * m.G + m.cg;
        const float4 d = __ldg(dz + i), x = __ldg(xhat + i);
        float4 o;
        o.x = ga.x * is.x * (d.x - inv_rows * (sd.x + x.x * sx.x));
        o.y = ga.y * is.y * (d.y - inv_rows * (sd.y + x.y * sx.y));
        o.z = ga.z * is.z * (d.z - inv_rows * (sd.z + x.z * sx.z));
        o.w = ga.w * is.w * (d.w - inv_rows * (sd.w + x.w * sx.w));
        dzpre[i] = o;
    }
}

// weight gradient, register-tiled: one CTA = a slab of rows x a 64 (outputs) x 32 (inputs) tile of dW; 64 rows at a
// time go through shared memory (128-bit loads, the NEXT block's loads are in flight in registers while the current one
// is multiplied); thread = 4 x 4 outputs over every S-th staged row, S = 256 / (4x4 tiles in this dW tile) row groups;
// the row groups are summed in shared memory, then one atomicAdd per output and CTA.  db[o] += sum_r dz[r, o] rides
// along (the i-tile 0 CTAs).  Needs Cout, c1, c2, ld1, ld2 multiples of 4 and 16-byte aligned pointers.
constexpr int kWgRows = 64, kWgTO = 64, kWgTI = 32;
__global__ void __launch_bounds__(kTrThreads)
train_wgrad4_kernel(const float *__restrict__ dz, int Cout, const float *__restrict__ in1, int ld1, int c1,
                    const float *__restrict__ in2, int ld2, int c2, long long rows, int rows_per_cta, float *__restrict__ dW,
                    float *__restrict__ db) {
    __shared__ __align__(16) float wg_smem[kWgRows * (kWgTO + 4) + kWgRows * (kWgTI + 4)];  // 6656 floats
    float (*sd)[kWgTO + 4] = reinterpret_cast<float (*)[kWgTO + 4]>(wg_smem);
    float (*sx)[kWgTI + 4] = reinterpret_cast<float (*)[kWgTI + 4]>(wg_smem + kWgRows * (kWgTO + 4));
    const int Kin = c1 + c2;
    const int o0 = blockIdx.y * kWgTO, i0 = blockIdx.z * kWgTI;
    const int to_n = min(kWgTO, Cout - o0) >> 2, ti_n = min(kWgTI, Kin - i0) >> 2;  // float4 groups in this tile
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    const int t = threadIdx.x;
    const int W = to_n * ti_n;                 // 4 x 4 register tiles in this dW tile (<= 128)
    int S = 1;                                 // row groups: the largest power of two with W * S <= 256, at most 64
    while (S < kWgRows && W * S * 2 <= kTrThreads) S *= 2;
    const int w = t % W, sgrp = t / W;
    const int ti = w % ti_n, to = w / ti_n;
    const bool worker = sgrp < S;
    const bool want_db = db != nullptr && blockIdx.z == 0;
    // staging map: thread t moves float4 #(t + 256 u) of the block; (row, group) fixed per u
    float4 pd[4], px[2];
    auto prefetch = [&](long long rb) {
        const int nrow = (int)min((long long)kWgRows, r1 - rb);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int q = t + kTrThreads * u, j = q / to_n, g = q - j * to_n;
            pd[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < kWgRows * to_n && j < nrow) pd[u] = __ldg(reinterpret_cast<const float4 *>(dz + (rb + j) * Cout + o0) + g);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int q = t + kTrThreads * u, j = q / ti_n, g = q - j * ti_n;
            px[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < kWgRows * ti_n && j < nrow) {
                const int ii = i0 + 4 * g;
                px[u] = ii < c1 ? __ldg(reinterpret_cast<const float4 *>(in1 + (rb + j) * ld1 + ii))
                                : __ldg(reinterpret_cast<const float4 *>(in2 + (rb + j) * ld2 + (ii - c1)));
            }
        }
    };
    float acc[4][4];
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
    if (r0 < r1) prefetch(r0);
    for (long long rb = r0; rb < r1; rb += kWgRows) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int q = t + kTrThreads * u, j = q / to_n, g = q - j * to_n;
            if (q < kWgRows * to_n) *reinterpret_cast<float4 *>(&sd[j][4 * g]) = pd[u];
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int q = t + kTrThreads * u, j = q / ti_n, g = q - j * ti_n;
            if (q < kWgRows * ti_n) *reinterpret_cast<float4 *>(&sx[j][4 * g]) = px[u];
        }
        __syncthreads();
        if (rb + kWgRows < r1) prefetch(rb + kWgRows);
        if (worker) {
#pragma unroll 4
            for (int j = sgrp; j < kWgRows; j += S) {
                const float4 d = *reinterpret_cast<const float4 *>(&sd[j][4 * to]);
                const float4 x = *reinterpret_cast<const float4 *>(&sx[j][4 * ti]);
                const float dv[4] = {d.x, d.y, d.z, d.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int a = 0; a < 4; a++) {
#pragma unroll
                    for (int b = 0; b < 4; b++) acc[a][b] = fmaf(dv[a], xv[b], acc[a][b]);
                    bsum[a] += dv[a];
                }
            }
        }
        __syncthreads();
    }
    // sum the S row groups through shared memory (the staging buffers are free now: S * W * 20 <= 5120 floats)
    float *red = wg_smem;
    if (worker) {
#pragma unroll
        for (int a = 0; a < 4; a++) {
#pragma unroll
            for (int b = 0; b < 4; b++) red[(sgrp * W + w) * 20 + 4 * a + b] = acc[a][b];
            red[(sgrp * W + w) * 20 + 16 + a] = bsum[a];
        }
    }
    __syncthreads();
    for (int q = t; q < W * 20; q += kTrThreads) {
        const int ww = q / 20, e = q - ww * 20;
        float v = 0.f;
        for (int g = 0; g < S; g++) v += red[(g * W + ww) * 20 + e];
        const int wti = ww % ti_n, wto = ww / ti_n;
        if (e < 16)
            atomicAdd(dW + (size_t)(o0 + 4 * wto + (e >> 2)) * Kin + i0 + 4 * wti + (e & 3), v);
        else if (want_db && wti == 0)
            atomicAdd(db + o0 + 4 * wto + (e - 16), v);
    }
}

// gather backward: dtable[rowidx[e], 4 + c] += dxf[e, c]
__global__ void __launch_bounds__(kTrThreads)
train_scatter_add_kernel(const float *__restrict__ dxf, const int *__restrict__ rowidx, long long edges, int Cin, int row_w,
                         float *__restrict__ dtable) {
    for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < edges * Cin; i += (long long)gridDim.x * kTrThreads) {
        const long long e = i / Cin;
        const int c = (int)(i % Cin);
        atomicAdd(dtable + (long long)rowidx[e] * row_w + 4 + c, dxf[i]);
    }
}

static bool col4_ok(int C, const void *a, const void *b = nullptr, const void *c = nullptr, const void *d = nullptr) {
    auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return C % 4 == 0 && C <= 1024 && al(a) && al(b) && al(c) && al(d);
}
// slabs of rows for the float4 kernels: about 4 CTAs per SM, at least one sweep of R rows each
static int col4_rows_per_cta(long long rows, int C) {
    const int R = kTrThreads / (C / 4);
    const long long per = (rows + 148 * 4 - 1) / (148 * 4);
    return (int)max((long long)max(R, 1) * 4, per);
}
static int tr_blocks(long long n) { return (int)max(1LL, min((n + kTrThreads - 1) / kTrThreads, 148LL * 16)); }

}  // namespace gg

using namespace gg;

extern "C" {

int gridgcn_train_edge_rows(const float *table, const int *nebidx, const float *cent, int B, int Nprev, int Cin, int O,
                            int K, int attfdim, float *xf, float *xa, int *rowidx, void *stream) {
    if (!table || !nebidx || !cent || !xf || !rowidx || B < 0 || Nprev < 1 || Cin < 0 || (Cin & 3) || O < 1 || K < 1)
        return GRIDGCN_EINVAL;
    if (attfdim > 0 && !xa) return GRIDGCN_EINVAL;
    if (B == 0) return 0;
    const int ain_p = attfdim > 0 ? (att_in_width(attfdim) + 3) / 4 * 4 : 0;
    const long long work = (long long)B * O * K * ((Cin > 0 ? Cin : 4) / 4);
    train_edge_rows_kernel<<<tr_blocks(work), kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        table, nebidx, reinterpret_cast<const float4 *>(cent), B, Nprev, Cin, O, K, attfdim, ain_p, xf, xa, rowidx);
    return (int)cudaGetLastError();
}

int gridgcn_train_col_sums(const float *a, const float *b, long long rows, int C, float *s0, float *s1, void *stream) {
    if (!a || !s0 || rows < 0 || C < 1) return GRIDGCN_EINVAL;
    if (rows == 0) return 0;
    if (col4_ok(C, a, b)) {
        const int per4 = col4_rows_per_cta(rows, C);
        const int nb = (int)((rows + per4 - 1) / per4);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const float4 *a4 = reinterpret_cast<const float4 *>(a), *b4 = reinterpret_cast<const float4 *>(b);
        if (b && s1) train_col_sums4_kernel<true, true><<<nb, kTrThreads, 0, st>>>(a4, b4, rows, C, per4, s0, s1);
        else if (b) train_col_sums4_kernel<true, false><<<nb, kTrThreads, 0, st>>>(a4, b4, rows, C, per4, s0, s1);
        else if (s1) train_col_sums4_kernel<false, true><<<nb, kTrThreads, 0, st>>>(a4, b4, rows, C, per4, s0, s1);
        else train_col_sums4_kernel<false, false><<<nb, kTrThreads, 0, st>>>(a4, b4, rows, C, per4, s0, s1);
        return (int)cudaGetLastError();
    }
    const int per = (int)max(64LL, (rows + 148 * 8 - 1) / (148 * 8));
    train_col_sums_kernel<<<(int)((rows + per - 1) / per), kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, b, rows, C, per, s0, s1);
    return (int)cudaGetLastError();
}

int gridgcn_train_bn_finalize(const float *s0, const float *s1, long long rows, int C, float eps, float momentum, float *mean,
                              float *invstd, float *running_mean, float *running_var, long long *num_batches_tracked,
                              void *stream) {
    if (!s0 || !s1 || !mean || !invstd || rows < 1 || C < 1) return GRIDGCN_EINVAL;
    train_bn_finalize_kernel<<<(C + kTrThreads - 1) / kTrThreads, kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        s0, s1, (float)rows, C, eps, momentum, mean, invstd, running_mean, running_var, num_batches_tracked);
    return (int)cudaGetLastError();
}

int gridgcn_train_bn_relu_fwd(const float *z, long long rows, int C, const float *mean, const float *invstd,
                              const float *gamma, const float *beta, float *y, void *stream) {
    if (!z || !mean || !invstd || !gamma || !beta || !y || rows < 0 || C < 1) return GRIDGCN_EINVAL;
    if (rows == 0) return 0;
    if (col4_ok(C, z, y) && col4_ok(C, mean, invstd, gamma, beta)) {
        const int per4 = col4_rows_per_cta(rows, C);
        train_bn_relu_fwd4_kernel<<<(int)((rows + per4 - 1) / per4), kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4 *>(z), rows, C, per4, reinterpret_cast<const float4 *>(mean),
            reinterpret_cast<const float4 *>(invstd), reinterpret_cast<const float4 *>(gamma),
            reinterpret_cast<const float4 *>(beta), reinterpret_cast<float4 *>(y));
        return (int)cudaGetLastError();
    }
    train_bn_relu_fwd_kernel<<<tr_blocks(rows * C), kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(z, rows * C, C, mean, invstd, gamma, beta, y);
    return (int)cudaGetLastError();
}

int gridgcn_train_relu_bwd_xhat(float *dy, const float *y, float *z, long long rows, int C, const float *mean,
                                const float *invstd, float *sum_dz, float *sum_dzx, void *stream) {
    if (!dy || !y || !z || !mean || !invstd || rows < 0 || C < 1) return GRIDGCN_EINVAL;
    if ((sum_dz == nullptr) != (sum_dzx == nullptr)) return GRIDGCN_EINVAL;
    if (rows == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (col4_ok(C, dy, y, z) && col4_ok(C, mean, invstd)) {
        const int per4 = col4_rows_per_cta(rows, C);
        train_relu_bwd_xhat4_kernel<<<(int)((rows + per4 - 1) / per4), kTrThreads, 0, st>>>(
            reinterpret_cast<float4 *>(dy), reinterpret_cast<const float4 *>(y), reinterpret_cast<float4 *>(z), rows, C, per4,
            reinterpret_cast<const float4 *>(mean), reinterpret_cast<const float4 *>(invstd), sum_dz, sum_dzx);
        return (int)cudaGetLastError();
    }
    train_relu_bwd_xhat_kernel<<<tr_blocks(rows * C), kTrThreads, 0, st>>>(dy, y, z, rows * C, C, mean, invstd);
    int rc = (int)cudaGetLastError();
    if (rc || !sum_dz) return rc;
    rc = gridgcn_train_col_sums(dy, nullptr, rows, C, sum_dz, nullptr, stream);
    return rc ? rc : gridgcn_train_col_sums(dy, z, rows, C, sum_dzx, nullptr, stream);
}

int gridgcn_train_bn_bwd(const float *dz, const float *xhat, long long rows, int C, const float *gamma, const float *invstd,
                         const float *sum_dz, const float *sum_dzx, float *dzpre, void *stream) {
    if (!dz || !xhat || !gamma || !invstd || !sum_dz || !sum_dzx || !dzpre || rows < 1 || C < 1) return GRIDGCN_EINVAL;
    if (col4_ok(C, dz, xhat, dzpre) && col4_ok(C, gamma, invstd, sum_dz, sum_dzx)) {
        const int per4 = col4_rows_per_cta(rows, C);
        train_bn_bwd4_kernel<<<(int)((rows + per4 - 1) / per4), kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float4 *>(dz), reinterpret_cast<const float4 *>(xhat), rows, C, per4, 1.0f / (float)rows,
            reinterpret_cast<const float4 *>(gamma), reinterpret_cast<const float4 *>(invstd),
            reinterpret_cast<const float4 *>(sum_dz), reinterpret_cast<const float4 *>(sum_dzx), reinterpret_cast<float4 *>(dzpre));
        return (int)cudaGetLastError();
    }
    train_bn_bwd_kernel<<<tr_blocks(rows * C), kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        dz, xhat, rows * C, C, 1.0f / (float)rows, gamma, invstd, sum_dz, sum_dzx, dzpre);
    return (int)cudaGetLastError();
}

int gridgcn_train_pool_fwd(const float *F, const float *A, long long centres, int K, int C, int pre_relu, const float *mask,
                           float *out, int ld_out, int *argmax, void *stream) {
    if (!F || !mask || !out || !argmax || centres < 0 || K < 1 || C < 1 || ld_out < C) return GRIDGCN_EINVAL;
    if (centres == 0) return 0;
    train_pool_fwd_kernel<<<tr_blocks(centres * C), kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(F, A, centres, K, C, pre_relu, mask, out, ld_out, argmax);
    return (int)cudaGetLastError();
}

int gridgcn_train_pool_bwd(const float *dout, int ld_out, const float *F, const float *A, const int *argmax, const float *mask,
                           long long centres, int K, int C, float *dF, float *dA, void *stream) {
    if (!dout || !F || !argmax || !mask || !dF || centres < 0 || K < 1 || C < 1) return GRIDGCN_EINVAL;
    if ((A == nullptr) != (dA == nullptr)) return GRIDGCN_EINVAL;
    if (centres == 0) return 0;
    train_pool_bwd_kernel<<<tr_blocks(centres * K * C), kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(dout, ld_out, F, A, argmax, mask, centres, K, C, dF, dA);
    return (int)cudaGetLastError();
}

int gridgcn_train_wgrad(const float *dz, int Cout, const float *in1, int ld1, int c1, const float *in2, int ld2, int c2,
                        long long rows, float *dW, float *db, void *stream) {
    if (!dz || !in1 || !dW || Cout < 1 || c1 < 1 || c2 < 0 || (c2 > 0 && !in2) || rows < 0) return GRIDGCN_EINVAL;
    if (rows == 0) return 0;
    auto al = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (Cout % 4 == 0 && c1 % 4 == 0 && c2 % 4 == 0 && ld1 % 4 == 0 && (c2 == 0 || ld2 % 4 == 0) && al(dz) && al(in1) && al(in2)) {
        const int tiles = ((Cout + kWgTO - 1) / kWgTO) * ((c1 + c2 + kWgTI - 1) / kWgTI);
        const long long want = max(1LL, (148LL * 3) / tiles);  // about 3 CTAs per SM over all tiles
        const int per4 = (int)max((long long)kWgRows, ((rows + want - 1) / want + kWgRows - 1) / kWgRows * kWgRows);
        dim3 grid4((unsigned)((rows + per4 - 1) / per4), (unsigned)((Cout + kWgTO - 1) / kWgTO), (unsigned)((c1 + c2 + kWgTI - 1) / kWgTI));
        train_wgrad4_kernel<<<grid4, kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(dz, Cout, in1, ld1, c1, in2, ld2, c2, rows, per4, dW, db);
        return (int)cudaGetLastError();
    }
    if (db) {
        const int rc = gridgcn_train_col_sums(dz, nullptr, rows, Cout, db, nullptr, stream);
        if (rc) return rc;
    }
    const int per = (int)max(256LL, (rows + 148 * 2 - 1) / (148 * 2));
    dim3 grid((unsigned)((rows + per - 1) / per), (unsigned)((Cout + 31) / 32), (unsigned)((c1 + c2 + 31) / 32));
    train_wgrad_kernel<<<grid, kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(dz, Cout, in1, ld1, c1, in2, ld2, c2, rows, per, dW);
    return (int)cudaGetLastError();
}

int gridgcn_train_scatter_add(const float *dxf, const int *rowidx, long long edges, int Cin, int row_w, float *dtable, void *stream) {
    if (!dxf || !rowidx || !dtable || edges < 0 || Cin < 1 || row_w < 4 + Cin) return GRIDGCN_EINVAL;
    if (edges == 0) return 0;
    train_scatter_add_kernel<<<tr_blocks(edges * Cin), kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(dxf, rowidx, edges, Cin, row_w, dtable);
    return (int)cudaGetLastError();
}

}  // extern "C"
