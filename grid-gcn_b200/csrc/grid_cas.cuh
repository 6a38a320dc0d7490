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

#include "cas_h_table.inc"

constexpr int kCasThreads = 256;

struct CasLayout {
    int cent_lin_out;  // [O]   centre voxel of each slot after sampling (read by the query kernel)
    int cent_acc_out;  // [4*O] its barycentre sums
    int cover;         // [ceil(Gp/2)] 16-bit voxel words when they do not fit in shared memory
};

// Voxel words live on a grid PADDED by r = (kernel-1)/2 cells on every side, so that the kernel^3
// neighbours of any in-grid voxel are plain offsets with no bounds test:
//   bit 15  set on the padding cells (never looks like count 0 or 1, never "occupied")
//   bit 14  the voxel is occupied (reference: coor_to_voxelidx >= 0)
//   bits 0-13  coverage count (<= kernel^3)
constexpr unsigned kCasPad = 0x8000u, kCasOcc = 0x4000u;

__host__ __device__ inline long long cas_padded_volume(const int grid[3], int ks) {
    const int r = (ks - 1) / 2;
    return (long long)(grid[0] + 2 * r) * (grid[1] + 2 * r) * (grid[2] + 2 * r);
}

__host__ inline size_t cas_smem_bytes(long long Gp, int O, bool cover_in_smem) {
    size_t b = (size_t)O * 8 + 16 + sizeof(unsigned short) * 2 * kCasHStates + 16 + 128;
    if (cover_in_smem) b += ((size_t)Gp * 2 + 15) & ~(size_t)15;
    return b;
}

template <bool COVER_SMEM>
__global__ void __launch_bounds__(kCasThreads)
cas_sampling_kernel(GridParams g, int *__restrict__ ws_base, WsLayout L, CasLayout C,
                    unsigned long long seed, float *__restrict__ centmsk, int *__restrict__ centnum) {
    extern __shared__ __align__(16) unsigned char cas_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const int O = g.O, ks = g.ks, S = ks * ks * ks, r = (ks - 1) / 2;
    int *ws = ws_base + (size_t)b * L.stride;
    const int M = ws[0];  // occupied voxels; the build kernel numbered them by first occurrence
    const int ncent = M < O ? M : O;
    const int *all_lin = ws + L.cent_lin;  // [M] linear voxel index by first-occurrence rank
    const float4 *all_acc = reinterpret_cast<const float4 *>(ws + L.cent_acc);

    int2 *slots = reinterpret_cast<int2 *>(cas_smem);  // per slot: padded voxel index, first-occurrence rank
    unsigned char *p = cas_smem + (((size_t)O * 8 + 15) & ~(size_t)15);
    // H state machine (cas_h_table.inc): valid for chains of up to kCasHSteps steps, i.e. kernel 1 or 3
    unsigned short *hnext = reinterpret_cast<unsigned short *>(p);
    p += (sizeof(unsigned short) * 2 * kCasHStates + 15) & ~(size_t)15;
    unsigned short *hbound = reinterpret_cast<unsigned short *>(p);
    p += 2 * 2 * 32;
    const bool one_chunk_table = S <= kCasHSteps;
    unsigned short *cover = COVER_SMEM ? reinterpret_cast<unsigned short *>(p)
                                       : reinterpret_cast<unsigned short *>(ws + C.cover);
    const int gx = g.grid[0], gy = g.grid[1], gz = g.grid[2], gxy = gx * gy;
    const int px = gx + 2 * r, py = gy + 2 * r, pz = gz + 2 * r, pxy = px * py;
    const int Gp = pxy * pz;
    auto padded_of = [&](int lin) {
        const int z = lin / gxy, y = (lin - z * gxy) / gx, x = lin - z * gxy - y * gx;
        return (z + r) * pxy + (y + r) * px + (x + r);
    };
    auto delta_of = [&](int k) {  // padded-index offset of neighbour k, raster order d -> h -> w
        return (k / (ks * ks) - r) * pxy + ((k % (ks * ks)) / ks - r) * px + (k % ks - r);
    };

    if (M > O) {
        for (int i = tid; i < 2 * kCasHStates; i += kCasThreads) hnext[i] = (&kCasHNext[0][0])[i];
        if (tid < 2 * (kCasHSteps + 1)) hbound[tid] = (&kCasHBound[0][0])[tid];
        // voxel words: padding everywhere, then the interior rows cleared, then the occupied flags
        unsigned *cover32 = reinterpret_cast<unsigned *>(cover);
        for (int i = tid; i < (Gp + 1) / 2; i += kCasThreads) cover32[i] = kCasPad | (kCasPad << 16);
        __syncthreads();
        for (int row = tid; row < gz * gy; row += kCasThreads) {
            const int z = row / gy, y = row - z * gy;
            unsigned short *dst = cover + (z + r) * pxy + (y + r) * px + r;
            for (int x = 0; x < gx; x++) dst[x] = 0;
        }
        __syncthreads();
        // (32-bit atomics on the 16-bit pairs: the flags and a count <= kernel^3 never carry over)
        for (int i = tid; i < M; i += kCasThreads) {
            const int q = padded_of(all_lin[i]);
            atomicOr(&cover32[q >> 1], kCasOcc << (16 * (q & 1)));
            if (i < O) slots[i] = make_int2(q, i);
        }
        __syncthreads();
        for (int i = tid; i < O * S; i += kCasThreads) {  // coverage counts of the initial incumbents
            const int o = i / S, q = slots[o].x + delta_of(i - o * S);
            atomicAdd(&cover32[q >> 1], 1u << (16 * (q & 1)));
        }
        __syncthreads();

        if (warp == 0 && one_chunk_table) {
            // Fast path (kernel 1 or 3: one lane per neighbour, H sums through the state machine).
            // Per batch of 32 challengers the padded index and the random slot are computed lane-parallel;
            // the challengers are then visited in order.  Per challenger: 2 shuffles, the (prefetched)
            // slot, one voxel-word load per lane and side, 2 ballots; when the value ranges of the two
            // sums overlap (kCasHBound) 2 more ballots and the two chains on lanes 0 / 1; on a swap the
            // -1 / +1 updates.
            const float Of = (float)O;
            const int dl = lane < S ? delta_of(lane) : 0;
            const unsigned short *hlo = hbound, *hhi = hbound + kCasHSteps + 1;
            for (int base = O; base < M; base += 32) {
                int my_q = 0, my_slot = 0;
                if (base + lane < M) {
                    my_q = padded_of(all_lin[base + lane]);
                    const float u = xorwow_first_uniform(seed + (unsigned long long)(base + lane));
                    my_slot = (int)(ceilf(__fmul_rn(Of, u)) - 1.0f);
                }
                const int cnt = min(32, M - base);
                int slot = __shfl_sync(kFull, my_slot, 0);
                int iq = slots[slot].x;
                for (int j = 0; j < cnt; j++) {
                    const int cq = __shfl_sync(kFull, my_q, j);
                    const int slot_n = __shfl_sync(kFull, my_slot, (j + 1) & 31);
                    int iq_n = slots[slot_n].x;  // prefetch; patched below if this challenger takes that slot
                    unsigned vi = kCasPad, vc = kCasPad;
                    if (lane < S) {
                        vi = cover[iq + dl];
                        vc = cover[cq + dl];
                    }
                    const bool m = (vi & ~kCasOcc) == 1u, a = (vc & ~kCasOcc) == 0u;
                    const unsigned ma = __ballot_sync(kFull, a), mr = __ballot_sync(kFull, m);
                    bool swap = false;
                    if (ma != 0u) {  // H_add == 0 can never exceed H_rmv >= 0
                        const int na = __popc(ma), nr = __popc(mr);
                        if (hlo[na] > hhi[nr]) {
                            swap = true;
                        } else if (hhi[na] > hlo[nr]) {  // ranges overlap: evaluate both sums exactly
                            const unsigned mao = __ballot_sync(kFull, a && (vc & kCasOcc));
                            const unsigned mro = __ballot_sync(kFull, m && (vi & kCasOcc));
                            unsigned todo = lane == 0 ? ma : (lane == 1 ? mr : 0u);
                            const unsigned occ = lane == 0 ? mao : mro;
                            int hs = 0;
                            while (todo) {
                                const unsigned bit = todo & (0u - todo);
                                todo ^= bit;
                                hs = hnext[((occ & bit) ? kCasHStates : 0) + hs];
                            }
                            swap = __shfl_sync(kFull, hs, 0) > __shfl_sync(kFull, hs, 1);
                        }
                    }
                    if (swap) {  // warp-uniform
                        if (lane == 0) slots[slot] = make_int2(cq, base + j);
                        if (slot_n == slot) iq_n = cq;
                        if (lane < S) cover[iq + dl] = (unsigned short)(vi - 1u);  // distinct half-words
                        __syncwarp();
                        if (lane < S) cover[cq + dl] += 1;
                        __syncwarp();
                    }
                    slot = slot_n;
                    iq = iq_n;
                }
            }
        } else if (warp == 0) {
            // General path (kernel >= 5): neighbours in chunks of 32, H in the reference's arithmetic.
            const float Of = (float)O;
            for (int base = O; base < M; base += 32) {
                int my_q = 0, my_slot = 0;
                if (base + lane < M) {
                    my_q = padded_of(all_lin[base + lane]);
                    const float u = xorwow_first_uniform(seed + (unsigned long long)(base + lane));
                    my_slot = (int)(ceilf(__fmul_rn(Of, u)) - 1.0f);
                }
                const int cnt = min(32, M - base);
                for (int j = 0; j < cnt; j++) {
                    const int cq = __shfl_sync(kFull, my_q, j);
                    const int slot = __shfl_sync(kFull, my_slot, j);
                    const int iq = slots[slot].x;
                    float H = 0.f;  // lane 0: H_add, lane 1: H_rmv
                    bool any_add = false;
                    for (int k0 = 0; k0 < S; k0 += 32) {
                        const int k = k0 + lane;
                        unsigned vi = kCasPad, vc = kCasPad;
                        if (k < S) {
                            const int dl = delta_of(k);
                            vi = cover[iq + dl];
                            vc = cover[cq + dl];
                        }
                        const bool m = (vi & ~kCasOcc) == 1u, a = (vc & ~kCasOcc) == 0u;
                        const unsigned ma = __ballot_sync(kFull, a);
                        const unsigned mao = __ballot_sync(kFull, a && (vc & kCasOcc));
                        const unsigned mr = __ballot_sync(kFull, m);
                        const unsigned mro = __ballot_sync(kFull, m && (vi & kCasOcc));
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
                    if (!any_add) continue;  // warp-uniform
                    if (__shfl_sync(kFull, H, 0) > __shfl_sync(kFull, H, 1)) {
                        if (lane == 0) slots[slot] = make_int2(cq, base + j);
                        for (int k = lane; k < S; k += 32) cover[iq + delta_of(k)] -= 1;
                        __syncwarp();
                        for (int k = lane; k < S; k += 32) cover[cq + delta_of(k)] += 1;
                    }
                    __syncwarp();
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
            const int cid = M > O ? slots[o].y : o;
            out_lin[o] = all_lin[cid];
            out_acc[o] = all_acc[cid];
        }
        centmsk[(size_t)b * O + o] = o < ncent ? 1.0f : 0.0f;
    }
    if (tid == 0) centnum[b] = ncent;
}

}  // namespace gg
