// gram_tc.cu -- segmentation-masked Gram matrices on the 5th-generation tensor cores.
//
// Replaces components/loss.py:96-102 (calculate_gram_matrix) for all K classes of one layer:
//     G_k = X_k^T X_k,   X_k = F * m_k  (rows = pixels scaled by the class mask, exactly the reference's formulation).
//
// GEMM view: D[c1, c2] = sum_px X[px, c1] X[px, c2]: the contraction runs over PIXELS while HBM holds pixel-major x
// channel, i.e. both operands are "MN-major" (the M/N index -- the channel -- is the contiguous one).  tcgen05 kind::f16
// accepts MN-major shared-memory operands (checked against a host reference by tests/cuda/umma_f16_mnmajor_probe.cu;
// kind::tf32 does not), so there is no on-chip transpose: one pipeline stage = a 2 x 16 pixel patch; TMA boxes
// (32 channels x 16 x 2 pixels) land pixel-major float32, and the transform warps -- which touch every element anyway
// for the mask scaling and the FP16 hi/lo split -- write the operand tiles in the canonical MN-major 128-byte-swizzle
// layout: one 128-byte row per pixel holding 64 channels, 8-pixel groups 1024 bytes apart, 64-channel groups 4096 bytes
// apart.  Diagonal tiles reuse the M-side operand for the N side.
// Only patches where the class mask is non-zero are visited (list built once per layer from the constant masks), so the
// work is ~(1 + boundary fraction) * 2 HW C^2 instead of K * 2 HW C^2.
//
// Precision: three FP16 terms per product with power-of-two scaling (max|F| from the producer's slot, max|mask| measured
// here), unbiased hi/lo splits and chunk promotion to registers, as in conv_tc.cu.
// Output: per-(class, split) partial tiles in the workspace; gram_reduce_kernel sums them in float64 (deterministic) and
// mirrors the upper triangle.
#include <cuda_fp16.h>

#include "tc_common.cuh"
#include "vgg.cuh"

namespace adpst {

constexpr int GM_PH = 2, GM_PW = 16, GM_PX = GM_PH * GM_PW;     // 32 pixels per stage = 2 UMMA K-steps of 16
constexpr int GM_BLK_BYTES = GM_PX * 128;                        // one landed box (32 channels x 32 pixels fp32): 4 KB
constexpr int GM_THREADS = 512;                                  // 4 warpgroups: {TMA, MMA, -, -}, transform x2, drain
constexpr int GM_REGS_CTRL = 40, GM_REGS_XFORM = 112, GM_REGS_DRAIN = 224;      // setmaxnreg budget, as in conv_tc.cu
constexpr int GM_CHUNK_ITERS = 4;                                // stages per promoted chunk (8 big MMAs)
constexpr int GM_RAW_STAGES_MAX = 8;                             // landed float32 tiles in flight: 4 (BN = 128) or 8 (BN = 64)
constexpr int GM_MASK_AHEAD = 3;                                 // mask values requested this many stages (per warpgroup) early
constexpr int GM_OP_STAGES = 2;                                  // FP16 operand tiles: slot g belongs to transform warpgroup g
constexpr int GM_GROUP_BYTES = (GM_PX / 8) * 1024;               // one 64-channel group of an operand tile: 4 KB

template <int BN> struct GramCfg {
    static constexpr int RAW_A = 4 * GM_BLK_BYTES;                // landed by TMA: 4 blocks of [32 px][32 ch]
    static constexpr int RAW_B = (BN / 32) * GM_BLK_BYTES;
    static constexpr int A_BYTES = 2 * GM_GROUP_BYTES;            // fp16 operand tile: 128 channels x 32 pixels
    static constexpr int B_BYTES = (BN / 64) * GM_GROUP_BYTES;
    // Two rings.  The kernel is bound by the latency of its TMA loads (the activations come from HBM: ~3000 clk against
    // ~400 clk of MMA work per stage), so the landed tiles get the deep ring; the operand tiles only need double buffering.
    // Both ring sizes are even, so with the two transform warpgroups alternating stages every slot always belongs to the
    // same warpgroup and it observes every phase of the slot's barriers (a parity wait cannot tell phases two apart).
    static constexpr int RAW_STAGES = BN == 64 ? 8 : 4;           // (freed as soon as they have been read)
    static constexpr int RAW_BYTES = BN == 64 ? 2 * GM_BLK_BYTES : RAW_A + RAW_B;   // C = 64: only two channel blocks exist
    static constexpr int OFF_RAW_B = RAW_A;
    static constexpr int OFF_AHI = 0, OFF_ALO = A_BYTES, OFF_BHI = 2 * A_BYTES, OFF_BLO = OFF_BHI + B_BYTES;
    static constexpr int OP_BYTES = OFF_BLO + B_BYTES;
    static constexpr int OFF_OP = RAW_STAGES * RAW_BYTES;
    static constexpr int OFF_BARS = OFF_OP + GM_OP_STAGES * OP_BYTES;
    static constexpr int SMEM_BYTES = OFF_BARS + 1024 + 256;
    static constexpr uint32_t TMEM_COLS = 4 * BN;
};

// shared-memory matrix descriptor, MN-major operand, 128-byte swizzle: LBO = distance between 64-element groups along
// M/N, SBO = distance between 8-row groups along K
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFF);
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}

template <int BN>
__global__ void __launch_bounds__(GM_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmF, const float* __restrict__ masks, const int* __restrict__ patch_ids,
               const int* __restrict__ patch_off, float* __restrict__ ws, int H, int W, int C, int splits, int tiles,
               int patches_w, const uint32_t* __restrict__ f_absmax, const uint32_t* __restrict__ m_absmax) {
    using Cfg = GramCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
    uint64_t* full = bars;                      // [RAW_STAGES]    raw tile landed
    uint64_t* raw_empty = bars + 8;             // [RAW_STAGES]    the transform warpgroup has read the raw tile
    uint64_t* ready = bars + 16;                // [GM_OP_STAGES]  operand tiles written
    uint64_t* empty = bars + 18;                // [GM_OP_STAGES]  the MMAs that read the operand tiles have retired
    uint64_t* chunk_full = bars + 20;
    uint64_t* chunk_empty = chunk_full + 2;
    uint64_t* small_full = chunk_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(small_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile pair (tm <= tn) in units of BN... the M side is always 128 channels, the N side BN channels
    int pair = blockIdx.x, tm = 0;
    while (pair >= tiles - tm) { pair -= tiles - tm; ++tm; }
    const int tn = tm + pair;
    const bool diag = (tm == tn);
    const int k = blockIdx.y / splits, sp = blockIdx.y - k * splits;
    const int p_lo = patch_off[k], p_hi = patch_off[k + 1];
    const int np = p_hi - p_lo;
    const int my_begin = p_lo + int((long long)np * sp / splits), my_end = p_lo + int((long long)np * (sp + 1) / splits);
    const int iters = my_end - my_begin;
    const int nchunks = (iters + GM_CHUNK_ITERS - 1) / GM_CHUNK_ITERS;
    const float* mk = masks ? masks + size_t(k) * H * W : nullptr;
    // operand scale: max|F * m| <= max|F| * max|m|
    int ex = tc::f16_scale_exponent(__ldg(f_absmax));
    if (m_absmax != nullptr) {
        const uint32_t mb = __ldg(m_absmax);
        if (mb > 0x3F800000u) ex -= int(mb >> 23) - 127 + 1;          // masks above 1 (not produced by loss.py, but allowed)
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::RAW_STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&raw_empty[s], 128);
        }
        for (int s = 0; s < GM_OP_STAGES; ++s) {
            tc::mbar_init(&ready[s], 128);
            tc::mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&chunk_full[b], 1);
            tc::mbar_init(&chunk_empty[b], 128);
        }
        tc::mbar_init(small_full, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmF);
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_small = tmem_base + 2 * BN;

    if (warp == 0) {
        // ================= TMA producer =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GM_REGS_CTRL));
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % Cfg::RAW_STAGES, round = it / Cfg::RAW_STAGES;
                tc::mbar_wait(&raw_empty[s], (round & 1) ^ 1);
                const int p = patch_ids[my_begin + it];
                const int y0 = (p / patches_w) * GM_PH, x0 = (p % patches_w) * GM_PW;
                uint8_t* st = smem + s * Cfg::RAW_BYTES;
                // BN == 64 <=> C == 64: only 64 of the M = 128 operand rows exist; rows 64..127 are never written and the
                // accumulator rows they produce are never read
                constexpr int a_blocks = BN == 64 ? 2 : 4;
                tc::mbar_arrive_expect_tx(&full[s], a_blocks * GM_BLK_BYTES + (diag ? 0 : Cfg::RAW_B));
#pragma unroll
                for (int b = 0; b < a_blocks; ++b)
                    tc::tma_load_4d(st + b * GM_BLK_BYTES, &tmF, &full[s], tm * 128 + b * 32, x0, y0, 0);
                if (!diag) {
#pragma unroll
                    for (int b = 0; b < BN / 32; ++b)
                        tc::tma_load_4d(st + Cfg::OFF_RAW_B + b * GM_BLK_BYTES, &tmF, &full[s], tn * BN + b * 32, x0, y0, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (warp-uniform loop, one elected lane issues; see conv_tc.cu) =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GM_REGS_CTRL));
        constexpr uint32_t idesc = tc::umma_idesc_f16(128, BN) | (1u << 15) | (1u << 16);      // A and B MN-major
        const uint32_t stage0 = tc::smem_u32(smem) + Cfg::OFF_OP;
        const uint64_t d_ahi = umma_desc_mnmajor_sw128(stage0 + Cfg::OFF_AHI, GM_GROUP_BYTES, 1024);
        const uint64_t d_alo = umma_desc_mnmajor_sw128(stage0 + Cfg::OFF_ALO, GM_GROUP_BYTES, 1024);
        const uint64_t d_bhi = diag ? d_ahi : umma_desc_mnmajor_sw128(stage0 + Cfg::OFF_BHI, GM_GROUP_BYTES, 1024);
        const uint64_t d_blo = diag ? d_alo : umma_desc_mnmajor_sw128(stage0 + Cfg::OFF_BLO, GM_GROUP_BYTES, 1024);
        for (int it = 0; it < iters; ++it) {
            const int s = it % GM_OP_STAGES, round = it / GM_OP_STAGES;
            const int c = it / GM_CHUNK_ITERS, cpos = it - c * GM_CHUNK_ITERS;
            const uint32_t tmem_big = tmem_base + uint32_t(c & 1) * BN;
            if (cpos == 0) tc::mbar_wait(&chunk_empty[c & 1], ((c >> 1) & 1) ^ 1);
            tc::mbar_wait(&ready[s], round & 1);
            tc::tcgen05_fence_after();
            const uint64_t soff = uint64_t(uint32_t(s) * uint32_t(Cfg::OP_BYTES >> 4));
            if (tc::elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < GM_PX / 16; ++ks) {                 // 16 pixels per MMA = two 8-pixel groups (2048 bytes)
                    const uint64_t koff = soff + uint64_t(ks * (2048 >> 4));
                    umma_f16_ss(tmem_small, d_alo + koff, d_bhi + koff, idesc, (it | ks) != 0);
                    umma_f16_ss(tmem_small, d_ahi + koff, d_blo + koff, idesc, 1);
                    umma_f16_ss(tmem_big, d_ahi + koff, d_bhi + koff, idesc, (cpos | ks) != 0);
                }
                tc::umma_commit(&empty[s]);
                if (cpos == GM_CHUNK_ITERS - 1 || it == iters - 1) tc::umma_commit(&chunk_full[c & 1]);
            }
            __syncwarp();
        }
        if (iters > 0 && tc::elect_one_sync()) tc::umma_commit(small_full);
        __syncwarp();
    } else if (warp < 4) {
        // (idle: completes the control warpgroup so that setmaxnreg applies to whole warpgroups)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GM_REGS_CTRL));
    } else if (warp < 12) {
        // ================= operand transform: X = m_k * F * 2^ex, split into FP16 hi / lo =================
        // Two warpgroups on alternating stages (one stage of this loop takes longer than its MMAs).
        // Thread t owns pixel p = t % 32 and the channel quarter cq = t / 32 of the patch: it reads one 128-byte row of
        // landed box cq and writes 64 bytes of the pixel's row in the hi tile and 64 bytes in the lo tile.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GM_REGS_XFORM));
        const int grp = (warp - 4) >> 2;
        const int t = (threadIdx.x - 128) & 127;                        // 0..127 within the warpgroup
        const int p = t & 31, cq = t >> 5;
        const float scale = tc::pow2f_int(ex);
        // operand tile address of (channel group cq / 2, pixel p): + ((unit ^ (p & 7)) << 4) for 16-byte unit `unit`
        const int orow = (cq >> 1) * GM_GROUP_BYTES + (p >> 3) * 1024 + (p & 7) * 128;
        const int u0 = (cq & 1) * 4;
        auto mask_of = [&](int it) -> float {                           // this pixel's mask value in the patch of stage `it`
            if (it >= iters) return 0.f;
            const int pid = patch_ids[my_begin + it];
            const int gy = (pid / patches_w) * GM_PH + p / GM_PW, gx = (pid % patches_w) * GM_PW + p % GM_PW;
            if (gy < H && gx < W) return mk ? __ldg(mk + size_t(gy) * W + gx) : 1.0f;
            return 0.f;
        };
        static_assert(Cfg::RAW_STAGES % 2 == 0 && Cfg::RAW_STAGES <= GM_RAW_STAGES_MAX && GM_OP_STAGES == 2, "every ring slot must always belong to the same warpgroup");
        // The mask value of a stage costs two DEPENDENT global loads (patch id, then the mask at that pixel): ~2 x 700 clk of
        // latency against ~700 clk of work per stage.  A queue of GM_MASK_AHEAD values per thread keeps that many of this
        // warpgroup's stages in flight (ncu before: long-scoreboard stalls 8.8 per issued instruction, tensor pipe 20 % busy).
        float mq[GM_MASK_AHEAD];
#pragma unroll
        for (int j = 0; j < GM_MASK_AHEAD; ++j) mq[j] = mask_of(grp + 2 * j);
        for (int it = grp; it < iters; it += 2) {
            // this warpgroup's stages: every second raw slot, operand slot grp
            const int rs = it % Cfg::RAW_STAGES, rround = it / Cfg::RAW_STAGES, os = grp, oround = it >> 1;
            const float sm = mq[0] * scale;
#pragma unroll
            for (int j = 0; j + 1 < GM_MASK_AHEAD; ++j) mq[j] = mq[j + 1];
            mq[GM_MASK_AHEAD - 1] = mask_of(it + 2 * GM_MASK_AHEAD);     // requested now, used GM_MASK_AHEAD stages of this group later
            tc::mbar_wait(&full[rs], rround & 1);                       // the raw tile has landed
            tc::mbar_wait(&empty[os], (oround & 1) ^ 1);                // the MMAs of stage it - 2 no longer read the operand slot
            const uint8_t* st = smem + rs * Cfg::RAW_BYTES;             // raw tile
            uint8_t* op = smem + Cfg::OFF_OP + os * Cfg::OP_BYTES;      // operand tiles
            if constexpr (BN == 64) {
                // C == 64: the 128 threads share 32 pixels x 64 channels, 16 channels (two 16-byte fp16 units) each
                const int e = cq;                                        // channels 16 e .. 16 e + 15
                const uint32_t raw = tc::smem_u32(st) + uint32_t((e >> 1) * GM_BLK_BYTES + p * 128);
                const uint32_t ohi = tc::smem_u32(op) + uint32_t(Cfg::OFF_AHI + (p >> 3) * 1024 + (p & 7) * 128);
                const uint32_t olo = tc::smem_u32(op) + uint32_t(Cfg::OFF_ALO + (p >> 3) * 1024 + (p & 7) * 128);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int c16 = (e & 1) * 4 + 2 * u;                 // 16-byte chunk of the landed 32-channel row
                    const float4 v0 = tc::lds128(raw + uint32_t((c16 ^ (p & 7)) << 4));
                    const float4 v1 = tc::lds128(raw + uint32_t(((c16 + 1) ^ (p & 7)) << 4));
                    const float x[8] = {v0.x * sm, v0.y * sm, v0.z * sm, v0.w * sm, v1.x * sm, v1.y * sm, v1.z * sm, v1.w * sm};
                    uint32_t hw[4], lw[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __half2 h = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
                        const float2 f = __half22float2(h);
                        const __half2 l = __floats2half2_rn((x[2 * j] - f.x) * 2048.0f, (x[2 * j + 1] - f.y) * 2048.0f);
                        hw[j] = *reinterpret_cast<const uint32_t*>(&h);
                        lw[j] = *reinterpret_cast<const uint32_t*>(&l);
                    }
                    const uint32_t chunk = uint32_t(((e * 2 + u) ^ (p & 7)) << 4);
                    tc::sts128(ohi + chunk, hw[0], hw[1], hw[2], hw[3]);
                    tc::sts128(olo + chunk, lw[0], lw[1], lw[2], lw[3]);
                }
                tc::mbar_arrive(&raw_empty[rs]);                         // (the stores above consumed every loaded value)
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(&ready[os]);
                continue;
            }
            const int nsets = diag ? 1 : 2;
            for (int set = 0; set < nsets; ++set) {
                if (set == 1 && cq * 32 >= BN) break;
                const uint32_t raw = tc::smem_u32(st) + uint32_t((set ? Cfg::OFF_RAW_B : 0) + cq * GM_BLK_BYTES + p * 128);
                const uint32_t ohi = tc::smem_u32(op) + uint32_t((set ? Cfg::OFF_BHI : Cfg::OFF_AHI) + orow);
                const uint32_t olo = tc::smem_u32(op) + uint32_t((set ? Cfg::OFF_BLO : Cfg::OFF_ALO) + orow);
#pragma unroll
                for (int u = 0; u < 4; ++u) {                            // 8 channels = one 16-byte unit of fp16
                    // landed layout: row = pixel (128 B), 16-byte chunk index XOR (row & 7)
                    const float4 v0 = tc::lds128(raw + uint32_t(((2 * u) ^ (p & 7)) << 4));
                    const float4 v1 = tc::lds128(raw + uint32_t(((2 * u + 1) ^ (p & 7)) << 4));
                    const float x[8] = {v0.x * sm, v0.y * sm, v0.z * sm, v0.w * sm, v1.x * sm, v1.y * sm, v1.z * sm, v1.w * sm};
                    uint32_t hw[4], lw[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __half2 h = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
                        const float2 f = __half22float2(h);
                        const __half2 l = __floats2half2_rn((x[2 * j] - f.x) * 2048.0f, (x[2 * j + 1] - f.y) * 2048.0f);
                        hw[j] = *reinterpret_cast<const uint32_t*>(&h);
                        lw[j] = *reinterpret_cast<const uint32_t*>(&l);
                    }
                    const uint32_t chunk = uint32_t(((u0 + u) ^ (p & 7)) << 4);
                    tc::sts128(ohi + chunk, hw[0], hw[1], hw[2], hw[3]);
                    tc::sts128(olo + chunk, lw[0], lw[1], lw[2], lw[3]);
                }
            }
            tc::mbar_arrive(&raw_empty[rs]);                             // (the stores above consumed every loaded value)
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&ready[os]);
        }
    } else {
        // ================= drain (chunk promotion) + store of the partial tile =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GM_REGS_DRAIN));
        const int q = warp & 3;
        const uint32_t lane_base = uint32_t(q * 32) << 16;
        float acc[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) acc[j] = 0.f;
        for (int c = 0; c < nchunks; ++c) {
            tc::mbar_wait(&chunk_full[c & 1], (c >> 1) & 1);
            tc::tcgen05_fence_after();
            const uint32_t src = tmem_base + uint32_t(c & 1) * BN + lane_base;
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tc::tmem_ld_32x32(src + c0, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
            }
            tc::tcgen05_fence_before();
            tc::mbar_arrive(&chunk_empty[c & 1]);
        }
        if (iters > 0) {
            tc::mbar_wait(small_full, 0);
            tc::tcgen05_fence_after();
        }
        const float inv_big = tc::pow2f_int(-2 * ex), inv_small = tc::pow2f_int(-2 * ex - 11);
        const int r = tm * 128 + q * 32 + lane;                         // Gram row (channel c1)
        float* out = ws + size_t(blockIdx.y) * C * C + size_t(r) * C + tn * BN;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            if (iters > 0) {
                tc::tmem_ld_32x32(tmem_small + lane_base + c0, v);
                tc::tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (q * 32 + lane < C - tm * 128) {                          // C == 64: rows 64..127 are duplicates
                float o[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = fmaf(__uint_as_float(v[j]), inv_small, acc[c0 + j] * inv_big);
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + c0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            }
        }
        tc::tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

bool gram_tc_eligible(int C) { return C == 64 || C % 128 == 0; }
int gram_tc_tiles(int C) { return C <= 128 ? 1 : C / 128; }

template <int BN>
static int launch_gram_tc_t(const CUtensorMap& tmF, const float* masks, const int* patch_ids, const int* patch_off, float* ws,
                            int H, int W, int C, int K, int splits, const uint32_t* f_absmax, const uint32_t* m_absmax,
                            cudaStream_t st) {
    using Cfg = GramCfg<BN>;
    auto kern = gram_tc_kernel<BN>;
    ADPST_ONCE_PER_DEVICE(ADPST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)));
    const int tiles = gram_tc_tiles(C);
    dim3 grid(tiles * (tiles + 1) / 2, K * splits);
    kern<<<grid, GM_THREADS, Cfg::SMEM_BYTES, st>>>(tmF, masks, patch_ids, patch_off, ws, H, W, C, splits, tiles,
                                                    (W + GM_PW - 1) / GM_PW, f_absmax, m_absmax);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

// partial Grams of all classes into ws[(k*splits + s)][C][C] (upper-triangular 128-tiles only).
// f_absmax: slot with max|F|; m_absmax: slot with max|mask| or NULL (all-ones mask).
int launch_gram_tc(const float* F, int H, int W, int C, const float* masks, int K, const int* patch_ids, const int* patch_off,
                   float* ws, int splits, const uint32_t* f_absmax, const uint32_t* m_absmax, cudaStream_t st) {
    CUtensorMap tmF;
    const uint64_t dims[4] = {uint64_t(C), uint64_t(W), uint64_t(H), 1};
    const uint64_t strides[3] = {uint64_t(C) * 4, uint64_t(W) * C * 4, uint64_t(H) * W * C * 4};
    const uint32_t box[4] = {32u, uint32_t(GM_PW), uint32_t(GM_PH), 1};
    int rc = tc::make_tensor_map_f32(&tmF, F, 4, dims, strides, box);
    if (rc != ADPST_OK) return rc;
    if (C == 64) return launch_gram_tc_t<64>(tmF, masks, patch_ids, patch_off, ws, H, W, C, K, splits, f_absmax, m_absmax, st);
    return launch_gram_tc_t<128>(tmF, masks, patch_ids, patch_off, ws, H, W, C, K, splits, f_absmax, m_absmax, st);
}

}  // namespace adpst
