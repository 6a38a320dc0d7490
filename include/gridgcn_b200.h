/*
 * gridgcn_b200.h -- C-ABI of libgridgcn_b200.so (sm_100a).
 *
 * These entry points are what the reference's operator plugin binds for the hot path named by
 * BASELINE.json:north_star: each one replaces the `Forward` of one MXNet operator registered by
 * gridifyop/additional.so (reference paths are relative to /root/reference/gridifyop/).
 * Plain pointers and sizes only; no torch / MXNet types.  See INTEGRATION.md for the
 * reference-side stub a maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the parameter name ends in `_host` or is one of the
 *    three-element parameter triples (coord_shift / voxel_size / grid_size), which are HOST arrays;
 *  - all tensors are dense, row-major, fp32 (`float`) or int32 (`int`), exactly the dtypes the
 *    reference infers (gridify-inl.h:203-213);
 *  - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream); all calls are
 *    asynchronous and re-entrant (no global state), the caller owns outputs AND workspace;
 *  - return value: 0 on success; a negative GRIDGCN_E* code for rejected arguments (nothing was
 *    launched); a positive value is a `cudaError_t` reported by a launch.  The library never
 *    aborts the process (the reference LOG(FATAL)s, gridify.cu:386).
 *
 * Canonical semantics: the reference kernels are non-deterministic
 * (atomics arrival order, time-seeded reservoirs).  These kernels produce the output of the
 * canonical schedule -- threads executed in ascending index, build before query, keep-first on
 * overflow -- which is the schedule the CPU oracle (oracle/gridgcn_oracle.c) restates.
 */
#ifndef GRIDGCN_B200_H_
#define GRIDGCN_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRIDGCN_ABI_VERSION 2   /* 2: gridgcn_mlp_t grew n_att_stages / localfdim / att_full */

/* rejected-argument codes (negative) */
#define GRIDGCN_EINVAL      (-1)  /* null pointer, negative size, even kernel_size, ...          */
#define GRIDGCN_ELIMIT      (-2)  /* outside the supported range (see each function)             */
#define GRIDGCN_EWORKSPACE  (-3)  /* workspace missing or smaller than *_workspace_bytes()       */

/* flags */
#define GRIDGCN_FLAG_DIST_FMA  1  /* d2 = fma(dz,dz,fma(dy,dy,dx*dx)) -- the contraction found in
                                     the reference's shipped sm_75 cubin -- instead of the
                                     canonical ((dx*dx+dy*dy)+dz*dz) of SURVEY.md s8c rule 10   */

#define GRIDGCN_FLAG_KNN_QUERY 2  /* gridgcn_gridify_occaware_fwd only: run the GridifyKNN query
                                     (K4) on the sampled centres instead of the Gridify one (K2) */

#define GRIDGCN_FLAG_STRICT_RESERVOIR 4  /* Gridify / Gridify_occaware (K2 query): reproduce the reference's
                                     reservoir over the candidates beyond max_p_grid (gridify.cu:259-270;
                                     its seed index_P*size+grid_pntidx is schedule independent) instead of
                                     the canonical keep-first rule.  cent.w is exact for integer-valued
                                     weights (the pipeline's: 1.0 inputs, neighbour counts afterwards).  */

int gridgcn_abi_version(void);

/* Human-readable text for a return code of this library (static storage). */
const char *gridgcn_strerror(int code);

/* ------------------------------------------------------------------------------------------ */
/* Gridify   -- replaces GridifyOp<gpu>::Forward, gridify-inl.h:99-128 + GridifyForward<gpu>,   */
/*              gridify.cu:294-413 (kernels gridify.cu:102-291).                                */
/* GridifyKNN-- replaces GridifyKNNOp<gpu>::Forward, gridifyknn-inl.h:99-128 + gridifyknn.cu    */
/*              :336-455 (kernels :115-333).  Same signature (the reference headers differ in   */
/*              identifiers only).                                                              */
/*                                                                                              */
/*  data      (B,N,4) f32  x,y,z,w           actual_numpoints (B) i32                           */
/*  nebidx    (B,O,P) i32  nebidxmsk (B,O,P) f32  cent (B,O,4) f32  centmsk (B,O) f32           */
/*  actual_centnum (B) i32          O = max_o_grid, P = max_p_grid  (gridify-inl.h:190-196)     */
/*  stride is accepted and ignored, as in the reference (gridify.cu:112, never read).           */
/*  Limits: kernel_size odd; P <= 128; grid_size[0]*[1]*[2] <= 262144; N < 2^24.                */
/* ------------------------------------------------------------------------------------------ */
size_t gridgcn_gridify_workspace_bytes(int B, int N, int max_o_grid, const int grid_size[3]);

int gridgcn_gridify_fwd(const float *data, const int *actual_numpoints, int B, int N,
                        int max_o_grid, int max_p_grid, int kernel_size, int stride, int loc,
                        const float coord_shift[3], const float voxel_size[3],
                        const int grid_size[3], int flags, int *nebidx, float *nebidxmsk,
                        float *cent, float *centmsk, int *actual_centnum, void *workspace,
                        size_t workspace_bytes, void *stream);

int gridgcn_gridify_knn_fwd(const float *data, const int *actual_numpoints, int B, int N,
                            int max_o_grid, int max_p_grid, int kernel_size, int stride, int loc,
                            const float coord_shift[3], const float voxel_size[3],
                            const int grid_size[3], int flags, int *nebidx, float *nebidxmsk,
                            float *cent, float *centmsk, int *actual_centnum, void *workspace,
                            size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------ */
/* Gridify_occaware -- Gridify with Coverage-Aware Sampling (CAS) of the centre voxels.         */
/*   Replaces the operators `Gridify_occaware` registered by gridifyop/additional.so, whose    */
/*   kernels gridify_kernel_build_index_occaware / gridify_occaware_sampling /                  */
/*   gridify_kernel_query_neighs_occaware exist in the reference ONLY as sm_61/sm_75 cubins    */
/*   (`cuobjdump -elf additional.so`; gridify_occaware.cu is not in the tree).    */
/*   Semantics: restated from the sm_75 SASS + arXiv:1912.02984 s3.2 under the canonical        */
/*   schedule (oracle/gridgcn_oracle.c, "Coverage-Aware Sampling"); parity unpinned.            */
/*   Same tensors as gridgcn_gridify_fwd.  `seed` replaces the reference's 2*tv_usec term: the   */
/*   challenger of first-occurrence rank i draws from XORWOW(seed + i).  Limits: those of        */
/*   Gridify, max_o_grid <= 8192, kernel_size <= 9.  flags: GRIDGCN_FLAG_KNN_QUERY, GRIDGCN_FLAG_DIST_FMA.     */
/* ------------------------------------------------------------------------------------------ */
size_t gridgcn_gridify_occaware_workspace_bytes(int B, int N, int max_o_grid, int kernel_size,
                                                const int grid_size[3]);

int gridgcn_gridify_occaware_fwd(const float *data, const int *actual_numpoints, int B, int N,
                                 int max_o_grid, int max_p_grid, int kernel_size, int stride,
                                 int loc, const float coord_shift[3], const float voxel_size[3],
                                 const int grid_size[3], int flags, unsigned long long seed,
                                 int *nebidx, float *nebidxmsk, float *cent, float *centmsk,
                                 int *actual_centnum, void *workspace, size_t workspace_bytes,
                                 void *stream);

/* ------------------------------------------------------------------------------------------ */
/* GridifyUp -- replaces GridifyUpOp<gpu>::Forward, gridify_up-inl.h:93-119 +                   */
/*              GridifyUpForward<gpu>, gridify_up.cu:228-324 (kernels :102-225).                */
/*  downdata (B,N,4) f32   updata (B,O,4) f32   down/up_actual_numpoints (B) i32                */
/*  nebidx (B,O,P) i32     nebidxmsk (B,O,P) f32                                               */
/* ------------------------------------------------------------------------------------------ */
size_t gridgcn_gridify_up_workspace_bytes(int B, int N, const int grid_size[3]);

int gridgcn_gridify_up_fwd(const float *downdata, const float *updata,
                           const int *down_actual_numpoints, const int *up_actual_numpoints,
                           int B, int N, int max_o_grid, int max_p_grid, int kernel_size,
                           const float coord_shift[3], const float voxel_size[3],
                           const int grid_size[3], int *nebidx, float *nebidxmsk,
                           void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------ */
/* KNN      -- replaces KNNForward<gpu>, k_nn-inl.h:94-113 (KNNKernel::Map :40-92,             */
/*             registration k_nn.cc:14-65).   Limit: k <= 128.                                  */
/* BallKNN  -- replaces BallKNNForward<gpu>, ball_k_nn-inl.h:97-116 (BallKNNKernel::Map        */
/*             :43-95, registration ball_k_nn.cc:14-65).   Limit: k <= 6 (best[6], :63-64).     */
/*  unknown (B,n,3) f32   known (B,m,3) f32   downnum, upnum (B) i32   idx (B,n,k) i32         */
/*  Rows >= upnum[b] are written as 0 (the reference leaves them unwritten, k_nn-inl.h:49-51). */
/* ------------------------------------------------------------------------------------------ */
int gridgcn_knn_fwd(const float *unknown, const float *known, const int *downnum,
                    const int *upnum, int B, int n, int m, int k, int flags, int *idx,
                    void *stream);

int gridgcn_ball_knn_fwd(const float *unknown, const float *known, const int *downnum,
                         const int *upnum, int B, int n, int m, int k, float radius, int flags,
                         int *idx, void *stream);

/* ------------------------------------------------------------------------------------------ */
/* GridConv -- one fused launch per layer that replaces batch_take_g (utils/ops.py:78-93) +     */
/*             sub_g_update (segmentation/models/gcn_module_g_att.py:172-287: geo features,     */
/*             verts_pair_func :120-170, aggregation_func :45-79 max pooling, update_func       */
/*             :24-43, centre mask :284-285), eval-mode BatchNorm folded into the 1x1 convs    */
/*             (utils/ops.py:149-158).                                                          */
/*                                                                                              */
/*  table   (B,Nprev,4+Cin) f32  rows [x y z w | Cin features] of the previous layer           */
/*                               (Cin = 0 for the first layer: has_feats=False)                 */
/*  nebidx  (B,O,K) i32          cent (B,O,4) f32       centmsk (B,O) f32                      */
/*  out     (B,O,4+Cout) f32     rows [cent x y z w | Cout features] = the next layer's table  */
/*                               (ggcn_models_g.py:186 concat(centers, center_feats))          */
/*  The MLP is described by `gridgcn_mlp_t`: folded weights W'(C_out,C_in) row-major and bias   */
/*  b'(C_out) per stage, feature chain first, then the attention stages.                        */
/*                                                                                              */
/*  Classification flavour of the block (classification/models/gcn_module_g.py:64-114,116-209):  */
/*  `localfdim` = 3 puts the geo vector in front of the gathered features (:186-191), the        */
/*  attention MLP has explicit widths (`n_att_stages` >= 2, the last one = C) and, with          */
/*  `att_full`, its stages after the first also see the feature MLP's output ("next", :91-93) or */
/*  its input ("last", :88-90).  These variants run in GRIDGCN_PRECISION_FP32 (one fused CUDA-   */
/*  core kernel) and GRIDGCN_PRECISION_TF32X3 (un-fused chain of tcgen05 row GEMMs over the edge  */
/*  rows: size the workspace with gridgcn_gridconv_edge_workspace_bytes); GRIDGCN_PRECISION_TF32  */
/*  returns GRIDGCN_ELIMIT for them.                                                             */
/* ------------------------------------------------------------------------------------------ */
#define GRIDGCN_MAX_STAGES 8

typedef struct {
    int n_feat_stages;                         /* len(pt_mlp_lst), 1..GRIDGCN_MAX_STAGES-2     */
    int attfdim;                               /* 10 (seg flavour) or 4 (dist, dxyz) or 0       */
    int feat_in;                               /* 3 (geo) when Cin==0, else Cin (+3 if localfdim)*/
    int widths[GRIDGCN_MAX_STAGES];            /* out width of feat stages, then of att stages  */
    const float *weight[GRIDGCN_MAX_STAGES];   /* device, (C_out, C_in) row-major, BN folded    */
    const float *bias[GRIDGCN_MAX_STAGES];     /* device, (C_out), BN folded                    */
    int pre_relu;                              /* configs["relu"], gcn_module_g_att.py:31-32    */
    int n_att_stages;                          /* 0 = 2 (segmentation block: C/4, C)            */
    int localfdim;                             /* 0, or 3: [geo_vec | features] as MLP input    */
    int att_full;                              /* GRIDGCN_ATT_FULL_*                            */
} gridgcn_mlp_t;

#define GRIDGCN_ATT_FULL_OFF  0
#define GRIDGCN_ATT_FULL_NEXT 1  /* attention stages >= 1 read [att_0 | feature MLP output]     */
#define GRIDGCN_ATT_FULL_LAST 2  /* attention stages >= 1 read [att_0 | feature MLP input]      */

#define GRIDGCN_PRECISION_FP32   0  /* CUDA-core fp32 FMA                                      */
#define GRIDGCN_PRECISION_TF32   1  /* tcgen05 kind::tf32, fp32 accumulate                     */
#define GRIDGCN_PRECISION_TF32X3 2  /* tcgen05 kind::tf32, 3-term error-compensated split      */

/* Tensor-core precisions need the weights re-laid out once per parameter set ("packed": hi/lo
 * tf32 operand images in tcgen05 shared-memory order) and, for layers with input features, a
 * workspace that receives the per-point transformed feature table. */
size_t gridgcn_gridconv_packed_bytes(const gridgcn_mlp_t *mlp_host, int Cin);
int gridgcn_gridconv_pack(const gridgcn_mlp_t *mlp_host, int Cin, void *packed, size_t packed_bytes,
                          void *stream);
size_t gridgcn_gridconv_workspace_bytes(const gridgcn_mlp_t *mlp_host, int B, int Nprev, int Cin);
/* Classification-block variants (localfdim / att_full / explicit attention widths) in GRIDGCN_PRECISION_TF32X3 run
 * as a chain of tensor-core row GEMMs over the edge rows (csrc/gridconv_cls_tc.cu) and need this much `workspace`
 * (a multiple of one cloud's worth, capped: the clouds are processed in chunks); 0 for the segmentation block. */
size_t gridgcn_gridconv_edge_workspace_bytes(const gridgcn_mlp_t *mlp_host, int B, int Cin, int O, int K);
/* GRIDGCN_PRECISION_FP32 keeps a tile's activations in shared memory; layers too wide for that (the
 * classification block's 256/512-channel layers) need this much `workspace` instead (0 otherwise). */
size_t gridgcn_gridconv_fp32_scratch_bytes(const gridgcn_mlp_t *mlp_host, int Cin, int K);

/* packed / workspace may be NULL for GRIDGCN_PRECISION_FP32. */
int gridgcn_gridconv_fwd(const float *table, const int *nebidx, const float *cent,
                         const float *centmsk, int B, int Nprev, int Cin, int O, int K,
                         const gridgcn_mlp_t *mlp_host, int precision, const void *packed,
                         void *workspace, size_t workspace_bytes, float *out, void *stream);

/* ------------------------------------------------------------------------------------------ */
/* Row MLP stage -- the per-centre 1x1 convolutions of the decoder half of sub_g_update           */
/* (segmentation/models/gcn_module_g_att.py:267-285 centre branch + concat, :24-43 update_func,   */
/* :284-285 centre mask) and of the segmentation head (ggcn_models_g.py:30-36), BN folded:        */
/*   out[r, 0:cout] = act_out( W * act_in([in1[r, 0:c1] | in2[r, 0:c2]]) + b ) * row_scale[r]     */
/* in1/in2/out are strided row views (ld = floats per row); in2 / row_scale may be NULL.  When     */
/* `cent` (rows,4) is given its rows are copied to out_table[r*ld_out + 0..3] (the [cent | feat]  */
/* table layout, ggcn_models_g.py:231).  Exact fp32 FMA on CUDA cores.                             */
/* ------------------------------------------------------------------------------------------ */
int gridgcn_rowmlp_fwd(const float *in1, int ld1, int c1, const float *in2, int ld2, int c2,
                       const float *weight, const float *bias, int cout, int relu_in, int relu_out,
                       const float *row_scale, float *out, int ld_out, const float *cent,
                       float *out_table, long long rows, void *stream);

/* The same operator on the tensor cores (tcgen05 kind::tf32, 3-pass hi/lo split: fp32-class accuracy, ~1e-6),
 * a persistent warp-specialised GEMM (csrc/rowgemm_tc.cu).  Views it cannot take (pointers / row strides / widths
 * not multiples of 16 bytes, cout > 256) run on the CUDA-core kernel above. */
int gridgcn_rowmlp_tc_fwd(const float *in1, int ld1, int c1, const float *in2, int ld2, int c2,
                          const float *weight, const float *bias, int cout, int relu_in, int relu_out,
                          const float *row_scale, float *out, int ld_out, const float *cent,
                          float *out_table, long long rows, void *stream);

/* ------------------------------------------------------------------------------------------ */
/* Training-mode GridConv block (BatchNorm with batch statistics, utils/ops.py:149-158 with       */
/* use_global_stats False; backward of gather / MLP / attention product / max pool): the kernels  */
/* around the tensor-core row GEMMs (gridgcn_rowmlp_tc_fwd computes Z = X W^T + b and dX = dZ W). */
/* All tensors are EDGE ROWS (edges = B*O*K rows, channels contiguous).  csrc/train_ops.cu;       */
/* orchestration: grid-gcn_b200/train_cuda.py.                                                    */
/* ------------------------------------------------------------------------------------------ */
/* xf (edges, Cin or 4 = [geo, 0]), xa (edges, pad4(att width)), rowidx (edges) = gathered table row */
int gridgcn_train_edge_rows(const float *table, const int *nebidx, const float *cent, int B, int Nprev, int Cin, int O,
                            int K, int attfdim, float *xf, float *xa, int *rowidx, void *stream);
/* s0[c] += sum_r a[r,c] * (b ? b[r,c] : 1);  s1[c] += sum_r a[r,c]^2 (s1 may be NULL); zero them first */
int gridgcn_train_col_sums(const float *a, const float *b, long long rows, int C, float *s0, float *s1, void *stream);
/* mean = s0 / rows, invstd = rsqrt(max(s1 / rows - mean^2, 0) + eps); moving statistics (may be NULL) updated as   */
/* running = (1 - momentum) * running + momentum * new, variance unbiased (torch.nn.BatchNorm convention);           */
/* *num_batches_tracked (int64, may be NULL) += 1                                                                    */
int gridgcn_train_bn_finalize(const float *s0, const float *s1, long long rows, int C, float eps, float momentum, float *mean,
                              float *invstd, float *running_mean, float *running_var, long long *num_batches_tracked,
                              void *stream);
int gridgcn_train_bn_relu_fwd(const float *z, long long rows, int C, const float *mean, const float *invstd,
                              const float *gamma, const float *beta, float *y, void *stream);
/* in place: dy <- dz = dy * (y > 0), z <- xhat = (z - mean) * invstd; when sum_dz / sum_dzx are given (both or    */
/* neither; zero them first) the same pass adds sum_r dz and sum_r dz * xhat per channel (BatchNorm backward)       */
int gridgcn_train_relu_bwd_xhat(float *dy, const float *y, float *z, long long rows, int C, const float *mean,
                                const float *invstd, float *sum_dz, float *sum_dzx, void *stream);
int gridgcn_train_bn_bwd(const float *dz, const float *xhat, long long rows, int C, const float *gamma, const float *invstd,
                         const float *sum_dz, const float *sum_dzx, float *dzpre, void *stream);
int gridgcn_train_pool_fwd(const float *F, const float *A, long long centres, int K, int C, int pre_relu, const float *mask,
                           float *out, int ld_out, int *argmax, void *stream);
int gridgcn_train_pool_bwd(const float *dout, int ld_out, const float *F, const float *A, const int *argmax, const float *mask,
                           long long centres, int K, int C, float *dF, float *dA, void *stream);
/* dW[o, i] += sum_r dz[r, o] * [in1 | in2][r, i];  db[o] += sum_r dz[r, o] (db may be NULL); zero them first */
int gridgcn_train_wgrad(const float *dz, int Cout, const float *in1, int ld1, int c1, const float *in2, int ld2, int c2,
                        long long rows, float *dW, float *db, void *stream);
int gridgcn_train_scatter_add(const float *dxf, const int *rowidx, long long edges, int Cin, int row_w, float *dtable, void *stream);

/* Self-test of the tcgen05 primitives (not an operator): D[128,N] = A[128,K] * B[N,K]^T on one CTA,
 * kind::tf32, nsplit 1 (plain) or 3 (error-compensated).  N % 16 == 0, N <= 256, K % 8 == 0. */
int gridgcn_debug_tc_gemm(const float *A, const float *B, float *D, int N, int K, int nsplit,
                          void *stream);

/* Debug / self-test: for each of n 64-bit seeds, the first uniform of this library's XORWOW (`ours`) and of
 * the device cuRAND the reference calls, curand_init(seed, 0, 0) + curand_uniform (`theirs`).  Not part of
 * the operator ABI. */
int gridgcn_debug_curand_first_uniform(const unsigned long long *seeds, int n, float *ours, float *theirs,
                                       void *stream);

/* Debug: see csrc/gridconv_fp32.cu.  8 x u64 device buffer of per-phase cycles, NULL = off. */
void gridgcn_debug_phase_buffer(unsigned long long *dev_buf);

#ifdef __cplusplus
}
#endif
#endif /* GRIDGCN_B200_H_ */
