// Hardware probe (not part of the library): tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY
// ("TS" form: tcgen05.mma [d], [a_tmem], b_desc, idesc, p).  Checks where the rows / k elements of A are
// expected (M = 128: row r in lane r, element k in column k; M = 64: row 16q+i in lane 32q+i, optionally
// offset by 16 lanes), that D may start at a column that is a multiple of 16 but not a power of two, and times
// back-to-back MMAs in the SS and TS forms for the shapes the first-layer GridConv kernel issues.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/ts_probe tools/ts_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../grid-gcn_b200/csrc/tc_common.cuh"
using namespace gg;

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

__host__ __device__ inline float a_val(int r, int k) { return (float)((r * 8 + k * 3) % 61 - 30); }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 3 + k * 5) % 17 - 8); }

// mode 0: M=128 TS.  mode 1: M=64 TS, A at lane offset a_off, D at lane offset d_off.  D columns start at dcol.
__global__ void __launch_bounds__(128) probe(float *D, long long *cyc, int mode, int a_off, int d_off, int dcol, int N) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = 8;
    const uint32_t lbo_b = (uint32_t)N * 16, lbo_a = 128 * 16;
    uint8_t *b = smem, *a_s = smem + 2 * lbo_b;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_init_fence(); }
    for (int e = tid; e < N * K; e += 128) {
        int r = e / K, k = e % K;
        *reinterpret_cast<float *>(b + tc::kmajor_off(r, k, lbo_b)) = b_val(r, k);
    }
    for (int e = tid; e < 128 * K; e += 128) {
        int r = e / K, k = e % K;
        *reinterpret_cast<float *>(a_s + tc::kmajor_off(r, k, lbo_a)) = a_val(r, k);
    }
    tc::fence_async_smem(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t a_col = 400;  // A lives in columns [400, 408)
    {   // every thread writes its lane of A (and zeroes of the D region first)
        float v[8];
        int row = -1;
        if (mode == 0) row = tid;
        else if ((lane >> 4) == (a_off >> 4)) row = 16 * warp + (lane & 15);
        for (int k = 0; k < 8; k++) v[k] = row >= 0 ? a_val(row, k) : 777.f;  // poison the lanes A should not use
        tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + a_col, v);
        tc::tmem_st_wait();
    }
    tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(mode == 0 ? 128 : 64, N);
        const uint32_t a_t = tmem + (mode == 0 ? 0u : ((uint32_t)a_off << 16)) + a_col;
        const uint32_t d_t = tmem + (mode == 0 ? 0u : ((uint32_t)d_off << 16)) + (uint32_t)dcol;
        mma_tf32_ts(d_t, a_t, tc::make_sdesc(tc::smem_u32(b), lbo_b), idesc, 0);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(dcol + c0), v);
        tc::tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[tid * 256 + c0 + j] = __uint_as_float(v[j]);
    }
    tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    // timing: 64 back-to-back accumulating MMAs, issue -> retired.  0: SS M=128  1: TS M=128  2: SS M=64  3: TS M=64
    for (int m = 0; m < 4; m++) {
        long long t0 = 0;
        if (tid == 0) {
            t0 = clock64();
            const int M = m < 2 ? 128 : 64;
            const uint32_t idesc = tc::make_idesc_tf32(M, N);
            const uint64_t bd = tc::make_sdesc(tc::smem_u32(b), lbo_b), ad = tc::make_sdesc(tc::smem_u32(a_s), lbo_a);
            for (int i = 0; i < 64; i++) {
                if (m & 1) mma_tf32_ts(tmem + (uint32_t)dcol, tmem + a_col, bd, idesc, 1);
                else tc::mma_tf32(tmem + (uint32_t)dcol, ad, bd, idesc, 1);
            }
            tc::mma_commit(&bar);
        }
        tc::mbar_wait(&bar, (m + 1) & 1);
        tc::fence_after_sync();
        if (tid == 0) cyc[m] = clock64() - t0;
        __syncthreads();
    }
    tc::fence_before_sync(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
    float *D; long long *cyc;
    cudaMallocManaged(&D, 128 * 256 * 4); cudaMallocManaged(&cyc, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    struct Cfg { int mode, a_off, d_off, dcol, N; };
    const Cfg cfgs[] = {{0, 0, 0, 0, 32}, {0, 0, 0, 48, 48}, {0, 0, 0, 96, 32}, {1, 0, 0, 0, 64}, {1, 16, 16, 64, 64},
                        {1, 16, 0, 128, 64}, {1, 0, 16, 256, 64}, {1, 0, 0, 0, 128}};
    for (const Cfg &c : cfgs) {
        cudaMemset(D, 0, 128 * 256 * 4);
        probe<<<1, 128, 2 * c.N * 16 + 2 * 128 * 16 + 1024>>>(D, cyc, c.mode, c.a_off, c.d_off, c.dcol, c.N);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode=%d a_off=%d d_off=%d dcol=%d N=%d: %s\n", c.mode, c.a_off, c.d_off, c.dcol, c.N, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        int bad = 0, checked = 0;
        const int M = c.mode == 0 ? 128 : 64;
        for (int r = 0; r < M; r++) {
            const int lane = c.mode == 0 ? r : (32 * (r / 16) + (r % 16) + c.d_off);
            for (int n = 0; n < c.N; n++) {
                float want = 0.f;
                for (int k = 0; k < 8; k++) want += a_val(r, k) * b_val(n, k);
                checked++;
                if (D[lane * 256 + n] != want) {
                    if (bad < 4) printf("  row %d (lane %d) col %d: got %g want %g\n", r, lane, n, D[lane * 256 + n], want);
                    bad++;
                }
            }
        }
        printf("  %d / %d mismatches;  64 MMAs: SS M=128 %lld, TS M=128 %lld, SS M=64 %lld, TS M=64 %lld cycles\n", bad, checked,
               cyc[0], cyc[1], cyc[2], cyc[3]);
    }
    return 0;
}
