// tc_common.cuh -- sm_100a building blocks shared by the tcgen05 kernels: mbarrier, TMA, TMEM, UMMA descriptors.
// Bit layouts follow the PTX ISA "tcgen05" chapter (shared-memory matrix descriptor, instruction descriptor).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace adpst {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// One lane of the (converged) warp is elected; the predicate is known to be warp-uniform-single, which lets the compiler
// keep UMMA descriptors in uniform registers instead of shuttling them per instruction.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}

// ---- explicit shared-memory accesses ----------------------------------------------------------------------------
// The kernels carve their dynamic shared memory up by hand (1024-byte alignment via integer arithmetic), which hides the
// address space from the compiler: plain dereferences become GENERIC loads/stores (LD.E / ST.E: longer latency, tracked
// on the long scoreboard).  The transform warps therefore address shared memory with 32-bit shared addresses.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- mbarrier ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Wait on an mbarrier phase.  The first poll is outside the loop: on the MMA-issuing thread every cycle spent here is a
// cycle the tensor pipe may idle.  After that the thread spins on try_wait (which itself suspends for a hardware-bounded
// time) and watches the WALL clock: %globaltimer_hi counts units of 2^32 ns (4.3 s) and is independent of SM clock
// throttling, preemption and time-slicing.  A wait that has made no progress for ADPST_MBAR_TIMEOUT_UNITS of them (default
// 7: 26-30 s of real time; a healthy stage takes microseconds) is a protocol bug and traps, which surfaces as a launch
// failure instead of a hung GPU.  One 32-bit register of state: the control warps of the tcgen05 kernels run with 40
// registers (setmaxnreg), and a 64-bit time stamp plus a poll counter in here made ptxas spill in their loops.
// Build with -DADPST_MBAR_TIMEOUT_UNITS=0 to spin without any bound.
#ifndef ADPST_MBAR_TIMEOUT_UNITS
#define ADPST_MBAR_TIMEOUT_UNITS 7
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ uint32_t global_timer_hi() {
    uint32_t t;
    asm volatile("mov.u32 %0, %%globaltimer_hi;" : "=r"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    if (mbar_try_wait(addr, parity)) return;
#if ADPST_MBAR_TIMEOUT_UNITS > 0
    const uint32_t t0 = global_timer_hi();
    while (!mbar_try_wait(addr, parity)) {
        if (global_timer_hi() - t0 > uint32_t(ADPST_MBAR_TIMEOUT_UNITS)) __trap();
    }
#else
    while (!mbar_try_wait(addr, parity)) {}
#endif
}

// ---- fences -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA (cp.async.bulk.tensor, tile mode) -------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- TMEM -------------------------------------------------------------------------------------------------------
// Whole-warp collectives (.sync.aligned).  ncols: power of two >= 32.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives lane (base_lane + t), columns c..c+31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// Store 32 consecutive 32-bit columns of this thread's TMEM lane (the mirror image of tmem_ld_32x32).
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
// Store 16 consecutive 32-bit columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA -------------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 B (32 fp32), 8-row groups 1024 B apart.
// bits [0,14) start>>4 | [16,30) LBO>>4 (ignored for swizzled K-major) | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFF);
    d |= uint64_t(1) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}
// (kind::tf32 helpers: used only by the probes under tests/cuda -- the product kernels run kind::f16)
// Instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major.
// [4,6) D format (1 = F32) | [7,10) A format (2 = TF32) | [10,13) B format | [15] A major | [16] B major | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T  -- issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with the A operand in tensor memory (lane = row of A, 8 consecutive 32-bit columns = the K slice): no shared-memory
// read for A.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Instruction descriptor for kind::f16 with both operands FP16 (format 0), fp32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// kind::f16, A operand in tensor memory: lane = row of A, 8 consecutive 32-bit columns = 16 packed fp16 (the K slice;
// element k sits in column k/2, bits 16*(k%2)..), B from shared memory.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All previously issued MMAs of this thread arrive on `bar` when they retire (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- FP16 operand splitting ("3xFP16": a*b = ah*bh + ah*bl + al*bh with fp32 accumulation) ------------------------
// A float32 tensor whose largest magnitude is known is mapped into FP16's range by a power-of-two scale s chosen so that
// 2^14 <= max|a| * s < 2^15.  Then  t = a*s (exact),  hi = fp16_rn(t) (11 significant bits, like TF32),
// lo = fp16_rn((t - hi) * 2^11) (t - hi is exact in fp32; the 2^11 keeps lo out of the subnormal range), so
// a*s = hi + lo * 2^-11 to 22 bits.  Values below 2^-29 of the tensor's maximum lose RELATIVE precision (fp16 subnormals)
// but their absolute error stays below 2^-50 of the maximum.  The accumulators are rescaled by powers of two (exact).
// `bits` = float bits of max|a| (non-negative floats order like unsigned integers, so the producers use atomicMax).
__host__ __device__ __forceinline__ int f16_scale_exponent(uint32_t absmax_bits) {
    const int E = int(absmax_bits >> 23) & 0xFF;          // max = 1.m * 2^(E-127)
    if (E == 0) return 0;                                   // all-zero (or subnormal) tensor: s = 1
    int e = 14 - (E - 127);
    return e > 60 ? 60 : (e < -60 ? -60 : e);              // |e_a + e_b| <= 120: one exact fp32 rescale at the end
}
__host__ __device__ __forceinline__ float pow2f_int(int e) {  // 2^e, |e| <= 126
#ifdef __CUDA_ARCH__
    return __uint_as_float(uint32_t(e + 127) << 23);
#else
    union { uint32_t u; float f; } c; c.u = uint32_t(e + 127) << 23; return c.f;
#endif
}

// ---- host: tensor maps --------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is resolved at run time (no link-time libcuda dependency: the library must load without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();
// fp32 tensor, `rank` dims (innermost first), byte strides for dims 1..rank-1, 128B swizzle, zero fill out of bounds.
int make_tensor_map_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box);
// same for an fp16 tensor (the box's innermost extent must be 64 elements = one 128-byte swizzle row)
int make_tensor_map_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box);

}  // namespace tc
}  // namespace adpst
