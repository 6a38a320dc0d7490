// Shared definitions of the fused GridConv kernels (sm_100a).
//
// One launch per layer replaces, for every (centre o, neighbour slot p) edge, the reference's
// op chain  batch_take_g (utils/ops.py:78-93) -> transpose -> sub_g_update
// (segmentation/models/gcn_module_g_att.py:172-287): geo_vec / geo_dist (:193-194), att_vec
// (:209-222), feature MLP + attention MLP + product (verts_pair_func :120-170), max pooling over
// the P slots (:57-59), pre-ReLU (update_func :31-32) and the centre mask (:284-285) -- about a
// dozen MXNet ops that each round-trip a (B,C,O,P) tensor through HBM.  Nothing but the output
// rows [cent | feats] is written here.
#pragma once
#include "../../include/gridgcn_b200.h"
#include "common.cuh"

namespace gg {

struct ConvParams {
    const float *table;    // (B*Nprev, 4+Cin)
    const int *nebidx;     // (B*O, K)
    const float4 *cent;    // (B*O)
    const float *centmsk;  // (B*O)
    float *out;            // (B*O, 4+Cout)
    float *scratch;        // fp32 path only: per-CTA activation buffers when they do not fit in shared memory
    int B, Nprev, Cin, O, K;
    int n_feat, attfdim, feat_in, pre_relu;
    int n_att, localfdim, att_full;      // attention stages (0 or >= 2), geo prefix width, GRIDGCN_ATT_FULL_*
    int n_stages;                        // n_feat + n_att
    int cin[GRIDGCN_MAX_STAGES];         // input width of each stage
    int cout[GRIDGCN_MAX_STAGES];        // output width of each stage
    const float *w[GRIDGCN_MAX_STAGES];  // (cout, cin) row-major, BN folded
    const float *bias[GRIDGCN_MAX_STAGES];
    int Cout;                            // = cout[n_feat-1]
};

// Row of the flattened table a neighbour index refers to.  mx.symbol.take's default mode is
// "clip" (utils/ops.py:92) applied AFTER the per-batch offset arange(B)*N was added (:90-91), so
// a BallKNN miss (-1) reads row b*N-1 and nothing ever reads out of bounds.
__device__ __forceinline__ long long take_row(int idx, int b, int Nprev, long long rows) {
    long long gi = (long long)idx + (long long)b * Nprev;
    return gi < 0 ? 0 : (gi >= rows ? rows - 1 : gi);
}

// att_vec layouts of sub_g_update (gcn_module_g_att.py:209-222).
//   attfdim 10: [dist, dx,dy,dz, cx,cy,cz, nx,ny,nz]   attfdim 4: [dist, dx,dy,dz]
//   attfdim <=3: [dx,dy,dz]
__device__ __forceinline__ void att_vector(int attfdim, float4 c, float nx, float ny, float nz,
                                           float *att /* >= 10 */, float &dx, float &dy, float &dz) {
    dx = nx - c.x;
    dy = ny - c.y;
    dz = nz - c.z;
    float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    if (attfdim <= 3) {
        att[0] = dx; att[1] = dy; att[2] = dz;
    } else {
        att[0] = dist; att[1] = dx; att[2] = dy; att[3] = dz;
        if (attfdim >= 10) {
            att[4] = c.x; att[5] = c.y; att[6] = c.z;
            att[7] = nx; att[8] = ny; att[9] = nz;
        }
    }
}

__host__ inline int att_in_width(int attfdim) { return attfdim <= 0 ? 0 : (attfdim <= 3 ? 3 : (attfdim < 10 ? 4 : 10)); }

}  // namespace gg
