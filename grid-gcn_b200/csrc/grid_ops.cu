// C-ABI entry points of the grid-query operators (see include/gridgcn_b200.h).
#include "../../include/gridgcn_b200.h"
#include "common.cuh"
#include "grid_build.cuh"
#include "grid_build_multi.cuh"
#include "grid_query.cuh"
#include "grid_cas.cuh"
#include "knn.cuh"

#include <algorithm>
#include <cstdlib>

namespace gg {

constexpr int kMaxGridVoxels = 262144;
constexpr size_t kBuildSmemLimit = 200 * 1024;

static int sm_count() {  // of the CURRENT device (the library may serve several devices of one process)
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    const bool cached = dev >= 0 && dev < 64;
    int n = cached ? cache[dev].load(std::memory_order_relaxed) : 0;
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        if (cached) cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

static int check_grid_args(int B, int N, int O, int P, int ks, const float *shift,
                           const float *voxel, const int *grid, GridParams &g) {
    if (B < 0 || N < 0 || O < 1 || P < 1 || ks < 1 || !shift || !voxel || !grid)
        return GRIDGCN_EINVAL;
    if ((ks & 1) == 0) return GRIDGCN_EINVAL;  // even kernel: coor_indx_b_origin unset, gridify.cu:248
    if (P > kMaxP || N >= (1 << 24) || (long long)B * O >= (1LL << 31)) return GRIDGCN_ELIMIT;  // 32-bit row arithmetic in the query kernels
    long long G = 1;
    for (int j = 0; j < 3; j++) {
        if (grid[j] < 1 || !(voxel[j] > 0.f)) return GRIDGCN_EINVAL;
        G *= grid[j];
        if (G > kMaxGridVoxels) return GRIDGCN_ELIMIT;
        g.shift[j] = shift[j];
        g.voxel[j] = voxel[j];
        g.grid[j] = grid[j];
        g.gridf[j] = (float)grid[j];
    }
    g.G = (int)G;
    g.W = (g.G + 31) / 32;
    g.B = B;
    g.N = N;
    g.O = O;
    g.P = P;
    g.ks = ks;
    return 0;
}

template <int THREADS>
static cudaError_t launch_build_t(const float *data, const int *npts, const GridParams &g,
                                  int *ws, const WsLayout &L, float *centmsk, int *centnum,
                                  int want_centers, cudaStream_t st) {
    bool in_smem = build_smem_words(g.N, g.G, true) * 4 <= kBuildSmemLimit;
    size_t smem = build_smem_words(g.N, g.G, in_smem) * 4;
    auto kern = grid_build_kernel<THREADS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kBuildSmemLimit);
    if (e != cudaSuccess) return e;
    kern<<<g.B, THREADS, smem, st>>>(reinterpret_cast<const float4 *>(data), npts, g, ws, L,
                                     centmsk, centnum, in_smem ? 1 : 0, want_centers);
    return cudaGetLastError();
}

// Point-parallel build (grid_build_multi.cuh): nine small launches over B * N threads.
static cudaError_t launch_build_multi(const float *data, const int *npts, const GridParams &g, int *ws,
                                      const WsLayout &L, float *centmsk, int *centnum, int want_centers,
                                      cudaStream_t st) {
    const float4 *d4 = reinterpret_cast<const float4 *>(data);
    auto blocks = [&](long long n) { return (int)std::max<long long>(1, std::min<long long>((n + kBmThreads - 1) / kBmThreads, (long long)sm_count() * 8)); };
    const long long BN = (long long)g.B * g.N;
    const int V = g.N < g.G ? g.N : g.G;
    bm_clear_kernel<<<blocks((long long)g.B * (g.W + (g.N + 31) / 32)), kBmThreads, 0, st>>>(g, ws, L);
    bm_voxelise_kernel<<<blocks(BN), kBmThreads, 0, st>>>(d4, npts, g, ws, L);
    bm_prefix_kernel<<<g.B, 1024, 0, st>>>(g, ws, L);
    bm_count_kernel<<<blocks(BN), kBmThreads, 0, st>>>(npts, g, ws, L);
    bm_scan_kernel<<<g.B, 1024, 0, st>>>(ws, L);
    bm_scatter_kernel<<<blocks(BN), kBmThreads, 0, st>>>(npts, g, ws, L);
    bm_rank_kernel<<<blocks(BN), kBmThreads, 0, st>>>(npts, g, ws, L, want_centers);
    if (!want_centers) bm_rank_heavy_kernel<<<g.B, 1024, 0, st>>>(npts, g, ws, L, 0);
    if (want_centers) {
        bm_firstpfx_kernel<<<g.B, 1024, 0, st>>>(npts, g, ws, L, centmsk, centnum);
        bm_centers_kernel<<<blocks((long long)g.B * V), kBmThreads, 0, st>>>(d4, g, ws, L);
    }
    return cudaGetLastError();
}

static bool build_multi_wanted(const GridParams &g) {
    // clouds that do not fit one CTA's shared memory, or too few clouds to occupy the SMs one CTA each
    static int force = -1;
    if (force < 0) {
        const char *e = getenv("GRIDGCN_BUILD_MULTI");
        force = e ? (e[0] == '1' ? 1 : (e[0] == '0' ? 0 : 2)) : 2;
    }
    if (force != 2) return force == 1;
    const bool in_smem = build_smem_words(g.N, g.G, true) * 4 <= kBuildSmemLimit;
    return !in_smem;
}

static cudaError_t launch_build(const float *data, const int *npts, const GridParams &g, int *ws,
                                const WsLayout &L, float *centmsk, int *centnum, int want_centers,
                                cudaStream_t st) {
    if (build_multi_wanted(g)) return launch_build_multi(data, npts, g, ws, L, centmsk, centnum, want_centers, st);
    if (g.N <= 2048)
        return launch_build_t<256>(data, npts, g, ws, L, centmsk, centnum, want_centers, st);
    return launch_build_t<kBuildThreads>(data, npts, g, ws, L, centmsk, centnum, want_centers, st);
}

static int query_grid_blocks(long long rows) {
    long long need = (rows + kQueryWarps - 1) / kQueryWarps;
    long long cap = (long long)sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

static int launch_gridify_query(bool knn, const float *data, const GridParams &g, const void *ws,
                                const WsLayout &L, int *nebidx, float *nebmsk, float *cent,
                                const int *centnum, cudaStream_t st) {
    const int B = g.B, N = g.N, O = g.O, P = g.P, ks = g.ks;
    const int blocks = query_grid_blocks((long long)B * O);
    const float4 *d4 = reinterpret_cast<const float4 *>(data);
    float4 *c4 = reinterpret_cast<float4 *>(cent);
    const int *w = static_cast<const int *>(ws);
    if (knn) {  // the sort key packs [voxel arrival order | point id] into 32 bits (grid_query.cuh)
        int combos = 0, vbits = 1;
        for (int l = 0; l < (ks + 1) / 2; l++) combos += (2 * l + 1) * (2 * l + 1) * (2 * l + 1);
        while ((1 << vbits) < combos) vbits++;
        if (N >= (1LL << (32 - vbits))) return GRIDGCN_ELIMIT;
    }
    if (!knn) {
        gridify_query_kernel<<<blocks, kQueryWarps * 32, 0, st>>>(d4, g, w, L, centnum, nebidx,
                                                                   nebmsk, c4);
    } else if (P <= 32) {
        gridify_knn_query_kernel<64><<<blocks, kQueryWarps * 32, 0, st>>>(d4, g, w, L, centnum,
                                                                           nebidx, nebmsk, c4);
    } else if (P <= 64) {
        gridify_knn_query_kernel<128><<<blocks, kQueryWarps * 32, 0, st>>>(d4, g, w, L, centnum,
                                                                            nebidx, nebmsk, c4);
    } else {
        gridify_knn_query_kernel<256><<<blocks, kQueryWarps * 32, 0, st>>>(d4, g, w, L, centnum,
                                                                            nebidx, nebmsk, c4);
    }
    return (int)cudaGetLastError();
}

static int gridify_impl(bool knn, const float *data, const int *npts, int B, int N, int O, int P,
                        int ks, int loc, const float *shift, const float *voxel, const int *grid,
                        int flags, int *nebidx, float *nebmsk, float *cent, float *centmsk,
                        int *centnum, void *ws, size_t ws_bytes, void *stream) {
    GridParams g{};
    int rc = check_grid_args(B, N, O, P, ks, shift, voxel, grid, g);
    if (rc) return rc;
    if (!data || !npts || !nebidx || !nebmsk || !cent || !centmsk || !centnum) return GRIDGCN_EINVAL;
    g.loc = loc;
    g.flags = flags;
    if (B == 0) return 0;
    if (!ws || ws_bytes < gridgcn_gridify_workspace_bytes(B, N, O, grid)) return GRIDGCN_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(ws) & 15) || (reinterpret_cast<uintptr_t>(data) & 15) ||
        (reinterpret_cast<uintptr_t>(cent) & 15))
        return GRIDGCN_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    WsLayout L = make_layout(N, O, g.G);
    cudaError_t e = launch_build(data, npts, g, static_cast<int *>(ws), L, centmsk, centnum, 1, st);
    if (e != cudaSuccess) return (int)e;
    return launch_gridify_query(knn, data, g, ws, L, nebidx, nebmsk, cent, centnum, st);
}

// Gridify_occaware: build (every occupied voxel numbered by first occurrence, with its barycentre)
// -> coverage-aware sampling of max_o of them (grid_cas.cuh) -> the Gridify / GridifyKNN query on the
// selected centres.  Workspace = the build layout sized for min(N, G) "centres" + the CAS extras.
struct CasPlan {
    WsLayout L;        // build layout (cent_lin / cent_acc hold every occupied voxel)
    WsLayout Lq;       // same memory, centre arrays redirected to the sampled ones (query kernel)
    CasLayout C;
    bool cover_in_smem;
    size_t smem;
};

static CasPlan make_cas_plan(int N, int O, const int grid[3], int ks) {
    CasPlan p;
    const int G = grid[0] * grid[1] * grid[2];
    const int V = N < G ? N : G;
    const long long Gp = cas_padded_volume(grid, ks);
    p.L = make_layout(N, V < 1 ? 1 : V, G);
    long long off = p.L.stride;
    p.C.cent_lin_out = (int)off;  off += round4(O);
    p.C.cent_acc_out = (int)off;  off += round4(4LL * O);
    p.cover_in_smem = cas_smem_bytes(Gp, O, true) <= kBuildSmemLimit;
    p.C.cover = (int)off;
    if (!p.cover_in_smem) off += round4((Gp + 1) / 2);
    p.L.stride = off;
    p.Lq = p.L;
    p.Lq.cent_lin = p.C.cent_lin_out;
    p.Lq.cent_acc = p.C.cent_acc_out;
    p.smem = cas_smem_bytes(Gp, O, p.cover_in_smem);
    return p;
}

static int gridify_occaware_impl(const float *data, const int *npts, int B, int N, int O, int P,
                                 int ks, int loc, const float *shift, const float *voxel,
                                 const int *grid, int flags, unsigned long long seed, int *nebidx,
                                 float *nebmsk, float *cent, float *centmsk, int *centnum, void *ws,
                                 size_t ws_bytes, void *stream) {
    GridParams g{};
    int rc = check_grid_args(B, N, O, P, ks, shift, voxel, grid, g);
    if (rc) return rc;
    if (!data || !npts || !nebidx || !nebmsk || !cent || !centmsk || !centnum) return GRIDGCN_EINVAL;
    if (O > 8192 || ks > 9) return GRIDGCN_ELIMIT;  // slot table in shared memory; padded grid
    g.loc = loc;
    g.flags = flags;
    if (B == 0) return 0;
    if (!ws || ws_bytes < gridgcn_gridify_occaware_workspace_bytes(B, N, O, ks, grid))
        return GRIDGCN_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(ws) & 15) || (reinterpret_cast<uintptr_t>(data) & 15) ||
        (reinterpret_cast<uintptr_t>(cent) & 15))
        return GRIDGCN_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const CasPlan p = make_cas_plan(N, O, grid, ks);
    GridParams gb = g;  // build: every occupied voxel is a candidate centre
    gb.O = N < g.G ? N : g.G;
    if (gb.O < 1) gb.O = 1;
    cudaError_t e = launch_build(data, npts, gb, static_cast<int *>(ws), p.L, nullptr, nullptr, 1, st);
    if (e != cudaSuccess) return (int)e;
    auto kern = p.cover_in_smem ? cas_sampling_kernel<true> : cas_sampling_kernel<false>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBuildSmemLimit);
    if (e != cudaSuccess) return (int)e;
    kern<<<B, kCasThreads, p.smem, st>>>(g, static_cast<int *>(ws), p.L, p.C, seed, centmsk, centnum);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    return launch_gridify_query((flags & GRIDGCN_FLAG_KNN_QUERY) != 0, data, g, ws, p.Lq, nebidx,
                                nebmsk, cent, centnum, st);
}

}  // namespace gg

using namespace gg;

extern "C" {

int gridgcn_abi_version(void) { return GRIDGCN_ABI_VERSION; }

const char *gridgcn_strerror(int code) {
    switch (code) {
        case 0: return "success";
        case GRIDGCN_EINVAL: return "invalid argument (null/misaligned pointer, negative size or even kernel_size)";
        case GRIDGCN_ELIMIT: return "argument outside the supported range";
        case GRIDGCN_EWORKSPACE: return "workspace missing or too small";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown gridgcn error";
    }
}

size_t gridgcn_gridify_workspace_bytes(int B, int N, int max_o_grid, const int grid_size[3]) {
    if (B <= 0 || N < 0 || max_o_grid < 1 || !grid_size) return 0;
    long long G = (long long)grid_size[0] * grid_size[1] * grid_size[2];
    if (G < 1 || G > kMaxGridVoxels) return 0;
    WsLayout L = make_layout(N, max_o_grid, (int)G);
    return (size_t)L.stride * 4 * (size_t)B;
}

size_t gridgcn_gridify_up_workspace_bytes(int B, int N, const int grid_size[3]) {
    return gridgcn_gridify_workspace_bytes(B, N, 1, grid_size);
}

int gridgcn_gridify_fwd(const float *data, const int *actual_numpoints, int B, int N,
                        int max_o_grid, int max_p_grid, int kernel_size, int stride, int loc,
                        const float coord_shift[3], const float voxel_size[3],
                        const int grid_size[3], int flags, int *nebidx, float *nebidxmsk,
                        float *cent, float *centmsk, int *actual_centnum, void *workspace,
                        size_t workspace_bytes, void *stream) {
    (void)stride;  // accepted and ignored, like the reference (gridify.cu:112)
    return gridify_impl(false, data, actual_numpoints, B, N, max_o_grid, max_p_grid, kernel_size,
                        loc, coord_shift, voxel_size, grid_size, flags, nebidx, nebidxmsk, cent,
                        centmsk, actual_centnum, workspace, workspace_bytes, stream);
}

int gridgcn_gridify_knn_fwd(const float *data, const int *actual_numpoints, int B, int N,
                            int max_o_grid, int max_p_grid, int kernel_size, int stride, int loc,
                            const float coord_shift[3], const float voxel_size[3],
                            const int grid_size[3], int flags, int *nebidx, float *nebidxmsk,
                            float *cent, float *centmsk, int *actual_centnum, void *workspace,
                            size_t workspace_bytes, void *stream) {
    (void)stride;
    return gridify_impl(true, data, actual_numpoints, B, N, max_o_grid, max_p_grid, kernel_size,
                        loc, coord_shift, voxel_size, grid_size, flags, nebidx, nebidxmsk, cent,
                        centmsk, actual_centnum, workspace, workspace_bytes, stream);
}

size_t gridgcn_gridify_occaware_workspace_bytes(int B, int N, int max_o_grid, int kernel_size,
                                                const int grid_size[3]) {
    if (B <= 0 || N < 0 || max_o_grid < 1 || kernel_size < 1 || kernel_size > 9 || !grid_size) return 0;
    long long G = (long long)grid_size[0] * grid_size[1] * grid_size[2];
    if (G < 1 || G > kMaxGridVoxels) return 0;
    return (size_t)make_cas_plan(N, max_o_grid, grid_size, kernel_size).L.stride * 4 * (size_t)B;
}

int gridgcn_gridify_occaware_fwd(const float *data, const int *actual_numpoints, int B, int N,
                                 int max_o_grid, int max_p_grid, int kernel_size, int stride,
                                 int loc, const float coord_shift[3], const float voxel_size[3],
                                 const int grid_size[3], int flags, unsigned long long seed,
                                 int *nebidx, float *nebidxmsk, float *cent, float *centmsk,
                                 int *actual_centnum, void *workspace, size_t workspace_bytes,
                                 void *stream) {
    (void)stride;
    return gridify_occaware_impl(data, actual_numpoints, B, N, max_o_grid, max_p_grid, kernel_size,
                                 loc, coord_shift, voxel_size, grid_size, flags, seed, nebidx,
                                 nebidxmsk, cent, centmsk, actual_centnum, workspace,
                                 workspace_bytes, stream);
}

int gridgcn_gridify_up_fwd(const float *downdata, const float *updata,
                           const int *down_actual_numpoints, const int *up_actual_numpoints,
                           int B, int N, int max_o_grid, int max_p_grid, int kernel_size,
                           const float coord_shift[3], const float voxel_size[3],
                           const int grid_size[3], int *nebidx, float *nebidxmsk,
                           void *workspace, size_t workspace_bytes, void *stream) {
    GridParams g{};
    int rc = check_grid_args(B, N, max_o_grid, max_p_grid, kernel_size, coord_shift, voxel_size,
                             grid_size, g);
    if (rc) return rc;
    if (!downdata || !updata || !down_actual_numpoints || !up_actual_numpoints || !nebidx ||
        !nebidxmsk)
        return GRIDGCN_EINVAL;
    if (B == 0) return 0;
    if (!workspace || workspace_bytes < gridgcn_gridify_up_workspace_bytes(B, N, grid_size))
        return GRIDGCN_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) || (reinterpret_cast<uintptr_t>(downdata) & 15) ||
        (reinterpret_cast<uintptr_t>(updata) & 15))
        return GRIDGCN_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the table is built over the DOWN points; O only sizes the (unused) centre arrays
    GridParams gb = g;
    gb.O = 1;
    gb.loc = 0;
    WsLayout L = make_layout(N, 1, g.G);
    cudaError_t e = launch_build(downdata, down_actual_numpoints, gb, static_cast<int *>(workspace),
                                 L, nullptr, nullptr, 0, st);
    if (e != cudaSuccess) return (int)e;
    const int blocks = query_grid_blocks((long long)B * max_o_grid);
    const float4 *u4 = reinterpret_cast<const float4 *>(updata);
    const int *w = static_cast<const int *>(workspace);
    const int P = max_p_grid;
    if (P <= 32)
        gridify_up_query_kernel<64><<<blocks, kQueryWarps * 32, 0, st>>>(u4, up_actual_numpoints, g,
                                                                          w, L, nebidx, nebidxmsk);
    else if (P <= 64)
        gridify_up_query_kernel<128><<<blocks, kQueryWarps * 32, 0, st>>>(u4, up_actual_numpoints, g,
                                                                           w, L, nebidx, nebidxmsk);
    else
        gridify_up_query_kernel<256><<<blocks, kQueryWarps * 32, 0, st>>>(u4, up_actual_numpoints, g,
                                                                           w, L, nebidx, nebidxmsk);
    return (int)cudaGetLastError();
}

static int knn_impl(bool ball, const float *unknown, const float *known, const int *downnum,
                    const int *upnum, int B, int n, int m, int k, float radius, int flags, int *idx,
                    void *stream) {
    if (B < 0 || n < 0 || m < 0 || k < 1) return GRIDGCN_EINVAL;
    if (!unknown || !known || !downnum || !upnum || !idx) return GRIDGCN_EINVAL;
    if (ball ? k > 6 : k > 128) return GRIDGCN_ELIMIT;  // best[6], ball_k_nn-inl.h:63-64
    if (B == 0 || n == 0) return 0;
    if (B > 65535) return GRIDGCN_ELIMIT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((n + kKnnThreads - 1) / kKnnThreads, B);
    const int fma = (flags & GRIDGCN_FLAG_DIST_FMA) ? 1 : 0;
    const float r2 = radius * radius;  // float product, ball_k_nn-inl.h:77
    if (ball)
        knn_kernel<8, true><<<grid, kKnnThreads, 0, st>>>(unknown, known, downnum, upnum, n, m, k,
                                                           r2, fma, idx);
    else if (k <= 8)
        knn_kernel<8, false><<<grid, kKnnThreads, 0, st>>>(unknown, known, downnum, upnum, n, m, k,
                                                            r2, fma, idx);
    else
        knn_kernel<128, false><<<grid, kKnnThreads, 0, st>>>(unknown, known, downnum, upnum, n, m,
                                                              k, r2, fma, idx);
    return (int)cudaGetLastError();
}

int gridgcn_knn_fwd(const float *unknown, const float *known, const int *downnum,
                    const int *upnum, int B, int n, int m, int k, int flags, int *idx,
                    void *stream) {
    return knn_impl(false, unknown, known, downnum, upnum, B, n, m, k, 0.f, flags, idx, stream);
}

int gridgcn_ball_knn_fwd(const float *unknown, const float *known, const int *downnum,
                         const int *upnum, int B, int n, int m, int k, float radius, int flags,
                         int *idx, void *stream) {
    return knn_impl(true, unknown, known, downnum, upnum, B, n, m, k, radius, flags, idx, stream);
}

}  // extern "C"
