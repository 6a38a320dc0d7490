// Hand-written tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX, no CUTLASS).
//
// Operand convention used by every tensor-core kernel in this library: both MMA operands are
// "K-major" fp32 matrices [rows x K] stored in shared memory in the NO-SWIZZLE canonical layout
//     element (r, k)  ->  base + (k/4)*LBO + (r/8)*SBO + (r%8)*16 + (k%4)*4        [bytes]
// i.e. 8-row x 16-byte core matrices; SBO = 128 B (the 16 row-groups of a 128-row operand are
// contiguous: one 2 KB "panel" per 4 consecutive k), LBO = panel stride.  One tcgen05.mma of
// kind::tf32 consumes K = 8 (two panels); the hardware reads fp32 bit patterns and uses the top 19
// bits (tf32).  Descriptor fields follow the PTX ISA "shared memory matrix descriptor" (version 1
// for sm_100) and "instruction descriptor" for kind::tf32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gg {
namespace tc {

constexpr uint32_t kSBO = 128;  // bytes between 8-row groups inside a panel

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// byte offset of element (r, k) inside an operand image with panel stride `lbo`
__device__ __forceinline__ uint32_t kmajor_off(uint32_t r, uint32_t k, uint32_t lbo) {
    return (k >> 2) * lbo + (r >> 3) * kSBO + (r & 7u) * 16u + (k & 3u) * 4u;
}

// shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version 1
// [46,48), base offset 0, swizzle NONE (bits 61-63 = 0)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) |
           ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46);
}

// instruction descriptor, kind::tf32: D fp32 (c_format 1 @4), A/B tf32 (format 2 @7, @10),
// both K-major (bits 15,16 = 0), N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, single-CTA, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// arrive on an mbarrier once every previously issued MMA of this thread has completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// plain arrive (count 1) by the calling thread
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// non-blocking phase test (polling loops that watch several barriers)
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: the warp SLEEPS in hardware until the phase completes (or ~1 ms pass)
// instead of spinning -- r02b profile: with the default (short) time limit the waiting warps of a
// warp-specialised kernel executed ~4000 polling instructions per 128-edge unit and starved the others.
__device__ __forceinline__ bool mbar_try_wait_sleep(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
        : "memory");
    return ok != 0;
}
// Bounded wait: a wrong descriptor must fail the launch (trap), never hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
#pragma unroll 1
    while (!mbar_try_wait_sleep(bar, parity))
        if (clock64() - t0 > 8000000000LL) __trap();  // ~4 s
}

// 1-D bulk copy global -> shared, completion counted on an mbarrier (TMA engine, SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// TMEM -> registers: this thread's lane, 32 consecutive fp32 columns starting at taddr's column
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// registers <-> TMEM, this thread's lane: 4 consecutive columns from / to 4 registers (TMEM as a
// per-thread stash).  tcgen05.st is asynchronous: tmem_st_wait() before the data is reused.
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr),
                 "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Round an fp32 value to tf32 (10-bit mantissa, low 13 bits zero), nearest / ties away.
__device__ __forceinline__ float rna_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}
// fp32 -> hi + lo, both exactly representable in tf32: hi = rna(v), lo = rna(v - hi) (v - hi is
// exact in fp32); |v - hi - lo| <= 2^-22 |v|.  Operands always reach the tensor core with their low
// 13 bits zero, so the result does not depend on how the hardware treats those bits.
__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
    hi = rna_tf32(v);
    lo = rna_tf32(v - hi);
}

// Cheaper split for the 3-pass mode, relying on the MEASURED behaviour of tcgen05 kind::tf32 on sm_100a:
// the tensor core TRUNCATES the low 13 mantissa bits of an fp32 operand (tools/tf32_probe.py: max
// |D - trunc-model| = 9e-7 vs 1.4e-2 for a round-to-nearest model; tests/test_gpu_tc.py pins it).
// So hi is the raw value (the hardware reads trunc(v)) and lo = v - trunc(v), exact in fp32; the hardware
// truncates lo to its top 11 bits, leaving |v - hi_eff - lo_eff| <= 2^-21 |v|.  2 instructions instead of 5.
template <int NSPLIT>
__device__ __forceinline__ void split_op(float v, float &hi, float &lo) {
    if (NSPLIT == 3) {
        hi = v;
        lo = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    } else {
        hi = rna_tf32(v);  // single pass: round to nearest halves the operand error
        lo = 0.f;
    }
}


// Packed fp32 pairs (sm_100a FADD2 / FMUL2: one instruction for two IEEE round-to-nearest results, bit-identical to
// the scalar operations).  Used by the epilogue roles, whose CUDA-core instruction count is what bounds them.
__device__ __forceinline__ void add2(float &x0, float &x1, float b0, float b1) {
    uint64_t a, b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(a));
}
// lo parts of the activation split (split_op<3>: lo = v - trunc_tf32(v)) of a pair: two LOP3 + one FADD2
__device__ __forceinline__ void split_lo2(float x0, float x1, float &lo0, float &lo1) {
    uint64_t a, b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "r"(__float_as_uint(x0) & 0xFFFFE000u), "r"(__float_as_uint(x1) & 0xFFFFE000u));
    asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo0), "=f"(lo1) : "l"(a));
}
__device__ __forceinline__ void mul2(float &x0, float &x1, float b0, float b1) {
    uint64_t a, b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(a));
}

}  // namespace tc
}  // namespace gg
