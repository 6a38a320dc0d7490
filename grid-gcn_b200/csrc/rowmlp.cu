// Row MLP stage (CUDA cores, exact fp32 FMA): out[r, :] = act( W * act_in([in1[r, :] | in2[r, :]]) + b ) * scale[r]
//
// The decoder half of the reference's GridConv block needs per-centre 1x1 convolutions that the
// encoder does not have: `mlp1d_c(center_ori_feats, center_dim)` on the up level's own [cent | feat]
// rows, the concat with the aggregated neighbour features, `update_func`'s pre-ReLU + `mlp1d_c(outDim)`
// and the centre mask (reference segmentation/models/gcn_module_g_att.py:267-285, :24-43), and the
// segmentation head (segmentation/models/ggcn_models_g.py:30-36).  Each is one call of this kernel
// with eval-mode BatchNorm folded into (W, b); inputs and outputs are strided row views so that the
// concat and the [cent | feat] table layout need no copy.
#include "../../include/gridgcn_b200.h"
#include "common.cuh"

namespace gg {

constexpr int kRmThreads = 256, kRmTile = 64, kRmK = 32;

__global__ void __launch_bounds__(kRmThreads)
rowmlp_kernel(const float *__restrict__ in1, int ld1, int c1, const float *__restrict__ in2, int ld2, int c2,
              const float *__restrict__ W, const float *__restrict__ bias, int cout, int relu_in,
              int relu_out, const float *__restrict__ scale, float *__restrict__ out, int ldo,
              const float *__restrict__ head4, float *__restrict__ out_head, long long rows) {
    __shared__ float As[kRmK][kRmTile + 4];  // [k][row]
    __shared__ float Ws[kRmK][kRmTile + 4];  // [k][col]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * kRmTile;
    const int col0 = blockIdx.y * kRmTile;
    const int K = c1 + c2;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += kRmK) {
        for (int e = tid; e < kRmTile * kRmK; e += kRmThreads) {
            const int r = e / kRmK, kk = e % kRmK, k = k0 + kk;
            float a = 0.f;
            if (row0 + r < rows && k < K) {
                a = k < c1 ? __ldg(in1 + (row0 + r) * ld1 + k) : __ldg(in2 + (row0 + r) * ld2 + (k - c1));
                if (relu_in) a = fmaxf(a, 0.f);
            }
            As[kk][r] = a;
            const int c = col0 + r;  // reuse r as the column index of the weight tile
            Ws[kk][r] = (c < cout && k < K) ? __ldg(W + (size_t)c * K + k) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < kRmK; kk++) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 w = *reinterpret_cast<const float4 *>(&Ws[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const long long r = row0 + ty * 4 + i;
        if (r >= rows) continue;
        const float s = scale ? __ldg(scale + r) : 1.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = col0 + tx * 4 + j;
            if (c < cout) {
                float v = acc[i][j] + __ldg(bias + c);
                if (relu_out) v = fmaxf(v, 0.f);
                out[r * ldo + c] = v * s;
            }
        }
        if (head4 && blockIdx.y == 0 && tx == 0) {  // copy the 4 centre columns of the output table
            const float4 h = __ldg(reinterpret_cast<const float4 *>(head4) + r);
            *reinterpret_cast<float4 *>(out_head + r * ldo) = h;
        }
    }
}

}  // namespace gg

extern "C" int gridgcn_rowmlp_fwd(const float *in1, int ld1, int c1, const float *in2, int ld2, int c2,
                                  const float *weight, const float *bias, int cout, int relu_in,
                                  int relu_out, const float *row_scale, float *out, int ld_out,
                                  const float *cent, float *out_table, long long rows, void *stream) {
    if (!in1 || !weight || !bias || !out || c1 < 1 || c2 < 0 || cout < 1 || rows < 0) return GRIDGCN_EINVAL;
    if (c2 > 0 && !in2) return GRIDGCN_EINVAL;
    if (ld1 < c1 || (c2 > 0 && ld2 < c2) || ld_out < cout) return GRIDGCN_EINVAL;
    if (cent && (!out_table || (reinterpret_cast<uintptr_t>(cent) & 15) ||
                 (reinterpret_cast<uintptr_t>(out_table) & 15) || (ld_out & 3)))
        return GRIDGCN_EINVAL;
    if (rows == 0) return 0;
    const long long tiles = (rows + gg::kRmTile - 1) / gg::kRmTile;
    if (tiles > 0x7fffffff) return GRIDGCN_ELIMIT;
    dim3 grid((unsigned)tiles, (unsigned)((cout + gg::kRmTile - 1) / gg::kRmTile));
    gg::rowmlp_kernel<<<grid, gg::kRmThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        in1, ld1, c1, in2, ld2, c2, weight, bias, cout, relu_in, relu_out, row_scale, out, ld_out, cent,
        out_table, rows);
    return (int)cudaGetLastError();
}
