// Coverage-Aware Sampling of the centre voxels (sm_100a) -- op Gridify_occaware.
//
// The reference ships this stage only as sm_61/sm_75 cubins inside gridifyop/additional.so
// (kernels gridify_kernel_build_index_occaware / gridify_occaware_sampling; their .cu is absent,
// SURVEY.md F3).  Semantics follow the SASS of those cubins and arXiv:1912.02984 s3.2, under the
// canonical schedule that oracle/gridgcn_oracle.c restates (parity unpinned):
//   incumbents  = the first max_o occupied voxels in first-occurrence order (slots 0..max_o-1);
//   cover[v]    = number of incumbents whose kernel^3 neighbourhood contains voxel v;
//   challengers = every later occupied voxel, in first-occurrence order, one attempt each:
//                 slot = ceilf(max_o * curand_uniform(XORWOW(seed + rank))) - 1,
//                 H_add = sum over the challenger's neighbours with cover == 0 of 0.7 (+0.3 if occupied),
//                 H_rmv = sum over the incumbent's  neighbours with cover == 1 of 0.7 (+0.3 if occupied),
//                 each += done as (float)((double)H + c) like the reference's F2F/DADD/F2F;
//                 swap iff H_add > H_rmv, then cover -= 1 around the incumbent, += 1 around the
//                 challenger.
// The reference runs one racing thread per challenger with atomicCAS on the slot and retries; the
// challengers depend on each other through `cover`, so here ONE WARP per cloud walks them in order
// with the 27 (kernel^3) neighbour lookups of both voxels spread over the lanes, the coverage
// counts (16 bit) and the occupancy bitmap in shared memory, and the two H chains evaluated on two
// lanes over ballot masks (only the neighbours that contribute are visited).
#pragma once
#include "common.cuh"

namespace gg {

constexpr int kCasThreads = 128;

struct CasLayout {
    int cent_lin_out;  // [O]   centre voxel of each slot after sampling (read by the query kernel)
    int cent_acc_out;  // [4*O] its barycentre sums
    int cover;         // [ceil(G/2)] 16-bit coverage counts when they do not fit in shared memory
};

__host__ inline size_t cas_smem_bytes(int G, int W, int O, bool cover_in_smem, bool bitmap_in_smem) {
    size_t b = (size_t)O * 20 + 16;
    if (cover_in_smem) b += ((size_t)G * 2 + 15) & ~(size_t)15;
    if (bitmap_in_smem) b += (size_t)W * 4;
    return b;
}

// curand_init(seed, 0, 0) followed by one curand_uniform (curand_kernel.h:772-798, 863-874).
__device__ __forceinline__ float xorwow_first_uniform(unsigned long long seed) {
    unsigned s0 = (unsigned)seed ^ 0xaad26b49u;
    unsigned s1 = (unsigned)(seed >> 32) ^ 0xf7dcefddu;
    unsigned t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    unsigned d = 6615241u + t1 + t0;
    unsigned v0 = 123456789u + t0, v4 = 5783321u + t0;
    unsigned t = v0 ^ (v0 >> 2);
    v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
    d += 362437u;
    return __fmaf_rn((float)(v4 + d), 2.3283064e-10f, 2.3283064e-10f / 2.0f);
}

__global__ void __launch_bounds__(kCasThreads)
cas_sampling_kernel(GridParams g, int *__restrict__ ws_base, WsLayout L, CasLayout C,
                    unsigned long long seed, float *__restrict__ centmsk, int *__restrict__ centnum,
                    int cover_in_smem, int bitmap_in_smem) {
    extern __shared__ __align__(16) unsigned char cas_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const int O = g.O, G = g.G, W = g.W, ks = g.ks, S = ks * ks * ks, r = (ks - 1) / 2;
    int *ws = ws_base + (size_t)b * L.stride;
    const int M = ws[0];  // occupied voxels; the build kernel numbered them by first occurrence
    const int ncent = M < O ? M : O;
    const int *all_lin = ws + L.cent_lin;  // [M] linear voxel index by first-occurrence rank
    const float4 *all_acc = reinterpret_cast<const float4 *>(ws + L.cent_acc);

    int4 *slot_zyxl = reinterpret_cast<int4 *>(cas_smem);  // z, y, x, linear index of each slot's voxel
    int *slot_cid = reinterpret_cast<int *>(slot_zyxl + O);  // its first-occurrence rank
    unsigned char *p = cas_smem + (((size_t)O * 20 + 15) & ~(size_t)15);
    unsigned short *cover;
    if (cover_in_smem) {
        cover = reinterpret_cast<unsigned short *>(p);
        p += ((size_t)G * 2 + 15) & ~(size_t)15;
    } else {
        cover = reinterpret_cast<unsigned short *>(ws + C.cover);
    }
    const unsigned *bitmap;
    if (bitmap_in_smem) {
        unsigned *bm = reinterpret_cast<unsigned *>(p);
        for (int i = tid; i < W; i += kCasThreads) bm[i] = (unsigned)ws[L.bitmap + i];
        bitmap = bm;
    } else {
        bitmap = reinterpret_cast<const unsigned *>(ws + L.bitmap);
    }
    const int gx = g.grid[0], gy = g.grid[1], gz = g.grid[2], gxy = gx * gy;

    if (M > O) {
        // coverage counts of the initial incumbents (32-bit atomics on the 16-bit pairs: a count never
        // exceeds min(O, S) < 65536, so no carry crosses into the neighbouring half-word)
        unsigned *cover32 = reinterpret_cast<unsigned *>(cover);
        for (int i = tid; i < (G + 1) / 2; i += kCasThreads) cover32[i] = 0u;
        for (int i = tid; i < O; i += kCasThreads) {
            const int lin = all_lin[i];
            const int z = lin / gxy, y = (lin - z * gxy) / gx;
            slot_zyxl[i] = make_int4(z, y, lin - z * gxy - y * gx, lin);
            slot_cid[i] = i;
        }
        __syncthreads();
        for (int i = tid; i < O * S; i += kCasThreads) {
            int o = i / S, k = i - o * S;
            const int4 c = slot_zyxl[o];
            int d = c.x + k / (ks * ks) - r, h = c.y + (k % (ks * ks)) / ks - r, w = c.z + k % ks - r;
            if (d < 0 || d >= gz || h < 0 || h >= gy || w < 0 || w >= gx) continue;
            int n = d * gxy + h * gx + w;
            atomicAdd(&cover32[n >> 1], (n & 1) ? 0x10000u : 1u);
        }
        __syncthreads();

        if (warp == 0) {
            // Lane-parallel per batch of 32 challengers: coordinates (one integer division pair per
            // lane instead of per challenger) and the random slot; then the challengers are visited in
            // order, each one costing a few shuffles, one slot load and the 2 x kernel^3 lookups.
            const float Of = (float)O;
            const bool one_chunk = S <= 32;
            const int od0 = lane / (ks * ks) - r, oh0 = (lane % (ks * ks)) / ks - r, ow0 = lane % ks - r;
            for (int base = O; base < M; base += 32) {
                int my_lin = 0, my_z = 0, my_y = 0, my_x = 0, my_slot = 0;
                if (base + lane < M) {
                    my_lin = all_lin[base + lane];
                    my_z = my_lin / gxy;
                    my_y = (my_lin - my_z * gxy) / gx;
                    my_x = my_lin - my_z * gxy - my_y * gx;
                    const float u = xorwow_first_uniform(seed + (unsigned long long)(base + lane));
                    my_slot = (int)(ceilf(__fmul_rn(Of, u)) - 1.0f);
                }
                const int cnt = min(32, M - base);
                for (int j = 0; j < cnt; j++) {
                    const int chal = __shfl_sync(kFull, my_lin, j);
                    const int cz = __shfl_sync(kFull, my_z, j), cy = __shfl_sync(kFull, my_y, j),
                              cx = __shfl_sync(kFull, my_x, j);
                    const int slot = __shfl_sync(kFull, my_slot, j);
                    const int4 inc4 = slot_zyxl[slot];  // z, y, x, linear index of the incumbent
                    const int iz = inc4.x, iy = inc4.y, ix = inc4.z;
                    float H = 0.f;  // lane 0: H_add, lane 1: H_rmv
                    bool any_add = false;
                    for (int k0 = 0; k0 < S; k0 += 32) {
                        const int k = k0 + lane;
                        int od = od0, oh = oh0, ow = ow0;
                        if (!one_chunk) {
                            od = k / (ks * ks) - r, oh = (k % (ks * ks)) / ks - r, ow = k % ks - r;
                        }
                        bool a = false, ao = false, m = false, mo = false;
                        if (k < S) {
                            int d = iz + od, h = iy + oh, w = ix + ow;
                            const bool in_i = d >= 0 && d < gz && h >= 0 && h < gy && w >= 0 && w < gx;
                            const int ni = in_i ? d * gxy + h * gx + w : 0;
                            d = cz + od, h = cy + oh, w = cx + ow;
                            const bool in_c = d >= 0 && d < gz && h >= 0 && h < gy && w >= 0 && w < gx;
                            const int nc = in_c ? d * gxy + h * gx + w : 0;
                            const unsigned short ci = cover[ni], cc = cover[nc];
                            const unsigned bi = bitmap[ni >> 5], bc = bitmap[nc >> 5];
                            m = in_i && ci == 1;
                            mo = m && ((bi >> (ni & 31)) & 1u);
                            a = in_c && cc == 0;
                            ao = a && ((bc >> (nc & 31)) & 1u);
                        }
                        const unsigned ma = __ballot_sync(kFull, a), mao = __ballot_sync(kFull, ao);
                        const unsigned mr = __ballot_sync(kFull, m), mro = __ballot_sync(kFull, mo);
                        any_add |= ma != 0u;
                        unsigned todo = lane == 0 ? ma : (lane == 1 ? mr : 0u);
                        const unsigned occ = lane == 0 ? mao : mro;
                        while (todo) {
                            const unsigned bit = todo & (0u - todo);
                            todo ^= bit;
                            H = (float)((double)H + 0.7);
                            if (occ & bit) H = (float)((double)H + 0.3);
                        }
                    }
                    if (!any_add) continue;  // H_add == 0 can never exceed H_rmv >= 0 (warp-uniform)
                    const float H_add = __shfl_sync(kFull, H, 0), H_rmv = __shfl_sync(kFull, H, 1);
                    if (H_add > H_rmv) {  // warp-uniform
                        if (lane == 0) {
                            slot_zyxl[slot] = make_int4(cz, cy, cx, chal);
                            slot_cid[slot] = base + j;
                        }
                        for (int k0 = 0; k0 < S; k0 += 32) {  // incumbent neighbourhood: -1
                            const int k = k0 + lane;
                            int od = od0, oh = oh0, ow = ow0;
                            if (!one_chunk) od = k / (ks * ks) - r, oh = (k % (ks * ks)) / ks - r, ow = k % ks - r;
                            const int d = iz + od, h = iy + oh, w = ix + ow;
                            if (k < S && d >= 0 && d < gz && h >= 0 && h < gy && w >= 0 && w < gx)
                                cover[d * gxy + h * gx + w] -= 1;  // distinct half-words per lane
                        }
                        __syncwarp();
                        for (int k0 = 0; k0 < S; k0 += 32) {  // challenger neighbourhood: +1
                            const int k = k0 + lane;
                            int od = od0, oh = oh0, ow = ow0;
                            if (!one_chunk) od = k / (ks * ks) - r, oh = (k % (ks * ks)) / ks - r, ow = k % ks - r;
                            const int d = cz + od, h = cy + oh, w = cx + ow;
                            if (k < S && d >= 0 && d < gz && h >= 0 && h < gy && w >= 0 && w < gx)
                                cover[d * gxy + h * gx + w] += 1;
                        }
                        __syncwarp();
                    }
                }
            }
        }
        __syncthreads();
    }

    // publish the selected centres for the query kernel, and the centre mask / count outputs
    int *out_lin = ws + C.cent_lin_out;
    float4 *out_acc = reinterpret_cast<float4 *>(ws + C.cent_acc_out);
    for (int o = tid; o < O; o += kCasThreads) {
        if (o < ncent) {
            const int cid = M > O ? slot_cid[o] : o;
            out_lin[o] = M > O ? slot_zyxl[o].w : all_lin[o];
            out_acc[o] = all_acc[cid];
        }
        centmsk[(size_t)b * O + o] = o < ncent ? 1.0f : 0.0f;
    }
    if (tid == 0) centnum[b] = ncent;
}

}  // namespace gg
