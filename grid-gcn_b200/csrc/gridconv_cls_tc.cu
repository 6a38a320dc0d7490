// Classification flavour of the GridConv block on the tensor cores -- GRIDGCN_PRECISION_TF32X3.
//
// Reference: classification/models/gcn_module_g.py:64-114 (verts_pair_func with att_full), :116-209
// (sub_g_update: localfdim = 3 puts the geo vector in front of the gathered features, explicit attention widths
// att_ele_dim, att_full "next" / "last" concatenates the feature MLP's output / input to the first attention stage's
// output), configs classification/configs/configs.yaml:47-68.  The fused tcgen05 kernels (gridconv_tc.cu,
// gridconv_first_ws.cu) implement the segmentation block only; r01 ran this flavour on the CUDA-core fp32 kernel
// alone (1.5e6 points/s on the shipped ModelNet40 ladder).
//
// Here the block runs UN-fused, every 1x1 convolution as one launch of the persistent tensor-core row GEMM
// (rowgemm_tc.cu) over the EDGE rows of a chunk of clouds:
//     cls_edge_rows_kernel   per edge: take (clip after the batch offset, utils/ops.py:90-92), geo = n - c, dist ->
//                            X_f = [geo, 0 | gathered feats]   (feature MLP input, 16-byte aligned layout)
//                            X_a = att_vec by attfdim           (padded to a multiple of 4)
//     feature MLP            row GEMM per stage  X_f -> ... -> F
//     attention MLP          row GEMM per stage; the concat of att_full is the GEMM's two-source input
//     cls_pool_kernel        max over the K slots of att * F, pre-ReLU, centre mask, [cent | feats] rows
// The activations travel through HBM between the launches (workspace: ~2.3 KB per edge on layer 0 of the shipped
// ladder), so the block is bandwidth bound, not tensor bound -- an order of magnitude above the CUDA-core kernel,
// not yet a fused pipeline.  Weight matrices whose input layout is padded (zero column after the geo vector,
// trailing zeros of att_vec) are re-laid-out once by gridgcn_gridconv_pack.
#include "gridconv_common.cuh"

#include <algorithm>

namespace gg {

int launch_rowgemm_tc(const float *in1, int ld1, int c1, const float *in2, int ld2, int c2, const float *W,
                      const float *bias, int N, int relu_in, int relu_out, const float *scale, float *out, int ldo,
                      const float *cent, float *out_table, long long rows, cudaStream_t st);  // rowgemm_tc.cu

struct ClsPlan {
    int fin_p, ain_p;            // padded widths of X_f and X_a
    int geo_in_feat;             // X_f starts with [geo, 0]
    long long w_f0, w_a0, w_a1;  // float offsets of the re-laid-out weights in `packed` (-1: the raw matrix is used)
    long long packed_floats;
    long long edge_floats;       // workspace floats per edge
};

static ClsPlan cls_plan(const ConvParams &p) {
    ClsPlan q{};
    q.geo_in_feat = (p.Cin == 0 || p.localfdim != 0) ? 1 : 0;
    q.fin_p = (q.geo_in_feat ? 4 : 0) + p.Cin;
    q.ain_p = p.attfdim > 0 ? (att_in_width(p.attfdim) + 3) / 4 * 4 : 0;
    long long off = 0;
    q.w_f0 = q.w_a0 = q.w_a1 = -1;
    if (q.geo_in_feat) { q.w_f0 = off; off += (long long)p.cout[0] * q.fin_p; }
    if (p.attfdim > 0) {
        const int a0 = p.n_feat;
        if (q.ain_p != p.cin[a0]) { q.w_a0 = off; off += (long long)p.cout[a0] * q.ain_p; }
        if (p.att_full == GRIDGCN_ATT_FULL_LAST && q.geo_in_feat) {
            q.w_a1 = off;
            off += (long long)p.cout[a0 + 1] * (p.cout[a0] + q.fin_p);
        }
    }
    q.packed_floats = std::max<long long>(off, 4);
    long long e = q.fin_p + q.ain_p;
    for (int s = 0; s < p.n_stages; s++) e += p.cout[s];
    q.edge_floats = e;
    return q;
}

bool cls_tc_ok(const ConvParams &p) {
    if ((p.Cin & 3) != 0) return false;
    for (int s = 0; s < p.n_stages; s++)
        if (p.cout[s] & 3) return false;
    return true;
}
long long cls_tc_packed_floats(const ConvParams &p) { return cls_plan(p).packed_floats; }
long long cls_tc_edge_bytes(const ConvParams &p) { return cls_plan(p).edge_floats * 4; }

// W (N x K) -> Wp (N x Kp): columns [0, gap_at) kept, gap_n zero columns inserted, the rest shifted; trailing zeros
__global__ void repack_cols_kernel(const float *__restrict__ W, float *__restrict__ Wp, int N, int K, int Kp, int gap_at,
                                   int gap_n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * Kp; i += gridDim.x * blockDim.x) {
        const int n = i / Kp, j = i % Kp;
        float v = 0.f;
        if (j < gap_at) v = j < K ? W[(size_t)n * K + j] : 0.f;
        else if (j >= gap_at + gap_n && j - gap_n < K) v = W[(size_t)n * K + j - gap_n];
        Wp[i] = v;
    }
}

int cls_tc_pack(const ConvParams &p, float *packed, cudaStream_t st) {
    const ClsPlan q = cls_plan(p);
    const int a0 = p.n_feat;
    auto run = [&](const float *W, long long off, int N, int K, int Kp, int gap_at, int gap_n) {
        repack_cols_kernel<<<std::min(1024, (N * Kp + 255) / 256), 256, 0, st>>>(W, packed + off, N, K, Kp, gap_at, gap_n);
    };
    if (q.w_f0 >= 0) run(p.w[0], q.w_f0, p.cout[0], p.cin[0], q.fin_p, 3, 1);            // [geo, 0 | feats]
    if (q.w_a0 >= 0) run(p.w[a0], q.w_a0, p.cout[a0], p.cin[a0], q.ain_p, p.cin[a0], q.ain_p - p.cin[a0]);
    if (q.w_a1 >= 0) run(p.w[a0 + 1], q.w_a1, p.cout[a0 + 1], p.cin[a0 + 1], p.cout[a0] + q.fin_p, p.cout[a0] + 3, 1);
    return (int)cudaGetLastError();
}

// One thread per (edge, 4-float group of the X_f row); group 0 also writes the X_a row.
__global__ void __launch_bounds__(256)
cls_edge_rows_kernel(ConvParams p, int b0, int nb, int fin_p, int ain_p, int geo_in_feat, float *__restrict__ xf,
                     float *__restrict__ xa) {
    const int groups = max(fin_p / 4, 1);
    const long long edges = (long long)nb * p.O * p.K;
    const long long rows_total = (long long)p.B * p.Nprev;
    const int row_w = 4 + p.Cin;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < edges * groups; i += (long long)gridDim.x * 256) {
        const long long e = i / groups;
        const int g = (int)(i % groups);
        const long long centre = (long long)b0 * p.O + e / p.K;  // global centre row
        const int b = (int)(centre / p.O);
        const long long row = take_row(__ldg(p.nebidx + centre * p.K + e % p.K), b, p.Nprev, rows_total);
        const float *src = p.table + row * row_w;
        if (g == 0) {
            const float4 c = __ldg(p.cent + centre);
            const float nx = __ldg(src), ny = __ldg(src + 1), nz = __ldg(src + 2);
            float att[12];
#pragma unroll
            for (int k = 0; k < 12; k++) att[k] = 0.f;
            float dx, dy, dz;
            att_vector(p.attfdim > 0 ? p.attfdim : 3, c, nx, ny, nz, att, dx, dy, dz);
            if (geo_in_feat) *reinterpret_cast<float4 *>(xf + e * fin_p) = make_float4(dx, dy, dz, 0.f);
            for (int k = 0; k < ain_p; k += 4)
                *reinterpret_cast<float4 *>(xa + e * ain_p + k) = make_float4(att[k], att[k + 1], att[k + 2], att[k + 3]);
            if (geo_in_feat) continue;
        }
        // feature columns: X_f[4g .. 4g+3] <- table row features (shifted by one group when the geo vector leads)
        const int fg = geo_in_feat ? g - 1 : g;
        *reinterpret_cast<float4 *>(xf + e * fin_p + 4 * g) = __ldg(reinterpret_cast<const float4 *>(src + 4 + 4 * fg));
    }
}

// One thread per (centre, channel): max over the K slots of att * F (gcn_module_g.py:108-112, :48-61), pre-ReLU, mask.
__global__ void __launch_bounds__(256)
cls_pool_kernel(ConvParams p, int b0, int nb, const float *__restrict__ F, const float *__restrict__ A) {
    const int C = p.Cout, out_w = 4 + C;
    const long long n = (long long)nb * p.O * C;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const long long lc = i / C;  // centre within the chunk
        const int ch = (int)(i % C);
        const float *f = F + lc * p.K * C + ch;
        const float *a = A ? A + lc * p.K * C + ch : nullptr;
        float m = -3.402823466e+38f;
        for (int k = 0; k < p.K; k++) {
            float v = __ldg(f + (size_t)k * C);
            if (a) v *= __ldg(a + (size_t)k * C);
            m = fmaxf(m, v);
        }
        const long long centre = (long long)b0 * p.O + lc;
        if (p.pre_relu) m = fmaxf(m, 0.f);
        p.out[centre * out_w + 4 + ch] = m * __ldg(p.centmsk + centre);
        if (ch < 4) p.out[centre * out_w + ch] = __ldg(reinterpret_cast<const float *>(p.cent + centre) + ch);
    }
}

static int gemm_cols(const float *in1, int ld1, int c1, const float *in2, int ld2, int c2, const float *W, const float *bias,
                     int N, float *out, long long rows, cudaStream_t st) {
    const int K = c1 + c2;
    for (int n0 = 0; n0 < N; n0 += 256) {
        const int rc = launch_rowgemm_tc(in1, ld1, c1, in2, ld2, c2, W + (size_t)n0 * K, bias + n0, std::min(256, N - n0), 0, 1,
                                         nullptr, out + n0, N, nullptr, nullptr, rows, st);
        if (rc != 0) return rc < 0 ? GRIDGCN_ELIMIT : rc;
    }
    return 0;
}

int launch_gridconv_cls_tc(const ConvParams &p, const float *packed, float *ws, size_t ws_bytes, cudaStream_t st) {
    if (!cls_tc_ok(p)) return GRIDGCN_ELIMIT;
    const ClsPlan q = cls_plan(p);
    const long long per_cloud = (long long)p.O * p.K * q.edge_floats * 4;
    const int nb_max = (int)std::min<long long>(p.B, (long long)(ws_bytes / (size_t)per_cloud));
    if (nb_max < 1 || !ws || !packed) return GRIDGCN_EWORKSPACE;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int a0 = p.n_feat;
    for (int b0 = 0; b0 < p.B; b0 += nb_max) {
        const int nb = std::min(nb_max, p.B - b0);
        const long long edges = (long long)nb * p.O * p.K;
        // workspace carve-up (floats): X_f | X_a | one buffer per stage output
        float *xf = ws, *xa = xf + edges * q.fin_p, *cur = xa + edges * q.ain_p;
        float *outb[GRIDGCN_MAX_STAGES];
        for (int s = 0; s < p.n_stages; s++) { outb[s] = cur; cur += edges * p.cout[s]; }
        const long long work = edges * std::max(q.fin_p / 4, 1);
        cls_edge_rows_kernel<<<(int)std::min<long long>((work + 255) / 256, (long long)sms * 16), 256, 0, st>>>(
            p, b0, nb, q.fin_p, q.ain_p, q.geo_in_feat, xf, xa);
        // feature MLP
        const float *src = xf;
        int ld = q.fin_p;
        for (int s = 0; s < p.n_feat; s++) {
            const float *W = (s == 0 && q.w_f0 >= 0) ? packed + q.w_f0 : p.w[s];
            const int rc = gemm_cols(src, ld, ld, nullptr, 0, 0, W, p.bias[s], p.cout[s], outb[s], edges, st);
            if (rc) return rc;
            src = outb[s];
            ld = p.cout[s];
        }
        const float *F = outb[p.n_feat - 1];
        const float *A = nullptr;
        if (p.attfdim > 0) {
            int rc = gemm_cols(xa, q.ain_p, q.ain_p, nullptr, 0, 0, q.w_a0 >= 0 ? packed + q.w_a0 : p.w[a0], p.bias[a0],
                               p.cout[a0], outb[a0], edges, st);
            if (rc) return rc;
            src = outb[a0];
            ld = p.cout[a0];
            for (int s = a0 + 1; s < p.n_stages; s++) {
                const float *in2 = nullptr;
                int ld2 = 0, c2 = 0;
                const float *W = p.w[s];
                if (s == a0 + 1 && p.att_full == GRIDGCN_ATT_FULL_NEXT) { in2 = F; ld2 = c2 = p.Cout; }
                if (s == a0 + 1 && p.att_full == GRIDGCN_ATT_FULL_LAST) {
                    in2 = xf; ld2 = c2 = q.fin_p;
                    if (q.w_a1 >= 0) W = packed + q.w_a1;
                }
                rc = gemm_cols(src, ld, ld, in2, ld2, c2, W, p.bias[s], p.cout[s], outb[s], edges, st);
                if (rc) return rc;
                src = outb[s];
                ld = p.cout[s];
            }
            A = outb[p.n_stages - 1];
        }
        const long long pw = (long long)nb * p.O * p.Cout;
        cls_pool_kernel<<<(int)std::min<long long>((pw + 255) / 256, (long long)sms * 16), 256, 0, st>>>(p, b0, nb, F, A);
    }
    return (int)cudaGetLastError();
}

}  // namespace gg
