// conv_tc.cu -- 3x3 SAME convolution (forward and data gradient) as an implicit GEMM on the 5th-generation tensor
// cores: TMA -> shared memory -> tcgen05.mma (kind::tf32) -> TMEM -> fused epilogue -> HBM.
//
// Replaces the Keras Conv2D calls behind components/VGG19/model.py:30 and their tape gradient (style_transfer.py:341).
//
// GEMM view per CTA:  D[128 pixels, BN channels] = sum over (tap, Cin chunk of 32)  A[128, 32] * B[BN, 32]^T
//   A: an 8 x 16 pixel tile of the NHWC activation, shifted by the tap; ONE 4-D TMA box (C=32, W=16, H=8, N=1) lands it
//      K-major with the 128-byte swizzle the UMMA descriptor expects, and TMA's out-of-bounds zero fill IS the SAME
//      padding, so there is no im2col buffer and no border code.
//   B: weights pre-arranged [tap][Cout][Cin] (K-major), 2-D TMA box (32, BN).
//
// Precision: the reference computes in float32 and parity is 1e-5.  One TF32 MMA (10-bit mantissa) is 1e-3.
// Every product is therefore formed as  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  ("3xTF32"), all three accumulating into the
// same fp32 TMEM accumulator:
//   a_hi = a rounded to TF32 with cvt.rna (exactly representable, so the tensor core's own fp32->tf32 conversion -- a
//          truncation, measured -- cannot change it),  a_lo = rna_tf32(a - a_hi)  (a - a_hi is exact in fp32).
//          Rounding instead of masking keeps both splits unbiased; a biased split compounds over the 12 layers.
//   b_hi / b_lo are split once when the weights are loaded; a_hi / a_lo are split by the "transform" warps between the
//   TMA landing and the MMA issue and handed to the tensor core through TENSOR MEMORY (tcgen05.st; the MMA's A operand
//   is a TMEM address), not shared memory.
// The dropped a_lo*b_lo term and the tf32 rounding of the lo parts are O(2^-22) relative.
//
// Warp roles (352 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer for the big term, warp 10 = MMA
// issuer for the small terms, warps 2..5 = operand transform (A hi/lo split into tensor memory), warps 6..9 = chunk
// promotion TMEM -> registers, then the epilogue (bias/ReLU or seed/mask -> global).  mbarrier rings: full (TMA landed) -> ready (A split done) -> empty (MMAs retired),
// chunk_full / chunk_empty for the two TMEM chunk buffers.
#include <stdlib.h>

#include "tc_common.cuh"
#include "vgg.cuh"

namespace adpst {
namespace tc {

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tensor_map_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(ADPST_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t d[5], s[5];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    static int promo = -1;
    if (promo < 0) { const char* ev = getenv("ADPST_TMA_PROMO"); promo = ev ? atoi(ev) : 2; }
    const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                                                                : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, cuuint32_t(rank), const_cast<void*>(base), d, s, b, e,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ADPST_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return ADPST_OK;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------------------------
// Accumulation accuracy.  Measured on B200 (scripts/tc_error_probe.py): tcgen05.mma adds into its fp32 accumulator with
// truncation, a systematic bias of about -2^-26 of the accumulator per MMA; over K = 4608 (1728 MMAs) that is -2.7e-5,
// above the 1e-5 parity bar.  Therefore:
//   * the big term a_hi*b_hi is accumulated in TMEM only over CHUNK_ITERS stages (8 MMAs) at a time, in two TMEM buffers
//     used alternately; four "drain" warps promote each finished chunk into fp32 REGISTERS with round-to-nearest adds
//     while the tensor core already works on the next chunk;
//   * the two small terms (2^-11 of the big one) accumulate in a third TMEM buffer for the whole tile -- their bias is
//     2^-11 smaller and negligible -- and are added once at the end.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TC_TH = 8, TC_TW = 16, TC_BM = TC_TH * TC_TW;     // 128 pixels = UMMA M
constexpr int TC_BK = 32;                                        // 32 fp32 = one 128-byte swizzle row
constexpr int TC_THREADS = 352;                                  // TMA, MMA(big), 4 transform, 4 drain/epilogue, MMA(small)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;                    // 16 KB
constexpr int TC_CHUNK_ITERS = 2;                                // stages per promoted chunk (8 big MMAs)
constexpr int TC_MAX_CLASSES = 32;                               // style mode: classes per launch (active set is a bit mask)

template <int BN> struct TcCfg {
    static constexpr int STAGES = 4;
    static constexpr int B_BYTES = BN * TC_BK * 4;
    static constexpr int STAGE_BYTES = TC_A_BYTES + 2 * B_BYTES;            // A (raw, as landed), B_hi, B_lo
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ +
                                      TC_MAX_CLASSES * TC_BM * 4 /*per-pixel class weights (style mode)*/;
    // tensor memory columns: big0 | big1 | small | A operand slots (2 x (hi: 32 columns, lo: 32 columns))
    static constexpr uint32_t COL_SMALL = 2 * BN, COL_A = 3 * BN, TMEM_COLS = 512;
};

// position of the n-th (0-based) set bit of m
__device__ __forceinline__ int nth_set_bit(uint32_t m, int n) {
    for (int i = 0; i < n; ++i) m &= m - 1;
    return __ffs(int(m)) - 1;
}

template <int BN, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                  const __grid_constant__ CUtensorMap tmBlo, const float* __restrict__ bias, float* __restrict__ Y,
                  const float* __restrict__ seed, const float* __restrict__ mask_src, int H, int W, int Cin, int Cout,
                  int tiles_w, const float* __restrict__ cls_masks, int num_cls, long long* __restrict__ dbg, int dbg_block, int dbg_flags) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    static_assert(BN == 64 || BN == 128, "tile width");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;                      // [STAGES]  A tile landed (the transform warps start on it at once)
    uint64_t* ready = bars + STAGES;            // [STAGES]  A split into hi/lo
    uint64_t* empty = bars + 2 * STAGES;        // [STAGES]  MMAs that read the stage have retired
    uint64_t* fullB = bars + 3 * STAGES;        // [STAGES]  (unused; kept so the barrier offsets stay put)
    uint64_t* chunk_full = bars + 4 * STAGES;   // [2]       a big-term chunk is complete in TMEM buffer b
    uint64_t* chunk_empty = chunk_full + 2;     // [2]       the drain warps have consumed TMEM buffer b
    uint64_t* small_full = chunk_empty + 2;     // [1]       every MMA of the tile has retired
    uint64_t* a_free = small_full + 1;          // [2]       the MMAs that read TMEM A slot j have retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + 2);
    float* cls_w = reinterpret_cast<float*>(bars + 32);                // [num_cls][128]  (style mode)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int y0 = (tile / tiles_w) * TC_TH, x0 = (tile % tiles_w) * TC_TW;
    const int n0 = blockIdx.y * BN;
    const int kchunks = Cin / TC_BK;
    // "taps" of the K loop: the 9 filter taps (convolution) or the classes whose mask is non-zero somewhere in this pixel
    // tile (style gradient: dF = sum_k w_k F D_k is a 1x1 convolution per class with a per-pixel weight w_k = m_k^2).
    uint32_t active = 0x1FFu;
    if (MODE == MODE_STYLE) {
        active = 0u;
        const int t = threadIdx.x;
        const int gy = y0 + t / TC_TW, gx = x0 + t % TC_TW;
        for (int k = 0; k < num_cls; ++k) {
            float w = 0.f;
            if (t < TC_BM) {
                if (gy < H && gx < W) {
                    w = cls_masks ? __ldg(cls_masks + size_t(k) * H * W + size_t(gy) * W + gx) : 1.0f;
                    w *= w;
                }
                cls_w[k * TC_BM + t] = w;
            }
            if (__syncthreads_or(w != 0.f)) active |= 1u << k;
        }
    }
    const int ntaps = __popc(active);
    const int iters = ntaps * kchunks;
    constexpr int chunk_iters = TC_CHUNK_ITERS;
    const int nchunks = (iters + chunk_iters - 1) / chunk_iters;
    // timeline instrumentation (development): dbg[role * 4096 + it * 2 + {0,1}] = clock64 for one chosen CTA
    const bool trace = dbg != nullptr && int(blockIdx.x) == dbg_block && blockIdx.y == 0;
    if (trace && threadIdx.x == 0) dbg[4 * 4096] = clock64();

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&ready[s], 128 + 1);      // 128 transform threads + the producer's expect_tx arrival (B bytes)
            tc::mbar_init(&empty[s], 2);            // both MMA-issuing warps commit
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&chunk_full[b], 1);
            tc::mbar_init(&chunk_empty[b], 128);
        }
        tc::mbar_init(small_full, 1);
        tc::mbar_init(&a_free[0], 2);
        tc::mbar_init(&a_free[1], 2);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmBhi);
        tc::tma_prefetch_desc(&tmBlo);
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_small = tmem_base + Cfg::COL_SMALL;
    const uint32_t tmem_a = tmem_base + Cfg::COL_A;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int s = 0, round = 0, slot = 0, kc = 0;
            for (int it = 0; it < iters; ++it) {
                tc::mbar_wait(&empty[s], (round & 1) ^ 1);
                const int tap = (MODE == MODE_STYLE) ? nth_set_bit(active, slot) : slot;
                const int kh = (MODE == MODE_STYLE) ? 1 : tap / 3, kw = (MODE == MODE_STYLE) ? 1 : tap - (tap / 3) * 3;
                uint8_t* st = smem + s * Cfg::STAGE_BYTES;
                if (dbg_flags & 8) tc::mbar_arrive(&full[s]);
                else {
                    tc::mbar_arrive_expect_tx(&full[s], TC_A_BYTES);
                    tc::tma_load_4d(st, &tmA, &full[s], kc * TC_BK, x0 + kw - 1, y0 + kh - 1, 0);
                }
                if (dbg_flags & 4) tc::mbar_arrive(&ready[s]);
                else {
                    tc::mbar_arrive_expect_tx(&ready[s], 2 * Cfg::B_BYTES);
                    tc::tma_load_2d(st + TC_A_BYTES, &tmBhi, &ready[s], kc * TC_BK, tap * Cout + n0);
                    tc::tma_load_2d(st + TC_A_BYTES + Cfg::B_BYTES, &tmBlo, &ready[s], kc * TC_BK, tap * Cout + n0);
                }
                if (++s == STAGES) { s = 0; ++round; }
                if (++kc == kchunks) { kc = 0; if (++slot == ntaps) slot = 0; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer, big term =================
        // Issuing one tcgen05.mma costs the issuing thread 40-48 clk (measured, tests/cuda/umma_queue_probe.cu) while a
        // 128x128x8 TF32 MMA executes in 64 clk, so ONE thread that also has ~240 clk of barrier work per stage cannot
        // keep the pipe fed with 12 MMAs per stage.  The issue is therefore split over two warps that own DIFFERENT
        // accumulators (no cross-thread ordering needed): this warp issues a_hi*b_hi into the chunk buffers, warp 10
        // issues the two small terms.  Both commit to empty[] / a_free[] (arrival count 2).
        // The whole warp runs the loop (warp-uniform control flow, descriptors in uniform registers); one elected lane
        // issues.  Descriptors are built once: per stage and K-step only the 14-bit address field changes.
        constexpr uint32_t idesc = tc::umma_idesc_tf32(TC_BM, BN);
        const uint32_t stage0 = tc::smem_u32(smem);
        const uint64_t d_bhi = tc::umma_desc_kmajor_sw128(stage0 + TC_A_BYTES, 1024);
        int s = 0, round = 0;
        for (int it = 0; it < iters; ++it) {
            const int c = it / chunk_iters, cpos = it - c * chunk_iters;
            const uint32_t tmem_big = tmem_base + uint32_t(c & 1) * BN;
            const uint32_t a_hi = tmem_a + uint32_t(it & 1) * 64;
            if (cpos == 0) {                                          // TMEM buffer must have been drained
                tc::mbar_wait(&chunk_empty[c & 1], ((c >> 1) & 1) ^ 1);
            }
            // ONE wait per stage: ready[s] completes when the B tiles have landed (transaction bytes) and the 128 transform
            // threads have stored A hi/lo into tensor memory.
            tc::mbar_wait(&ready[s], round & 1);
            tc::tcgen05_fence_after();
            const uint64_t soff = uint64_t(uint32_t(s) * uint32_t(Cfg::STAGE_BYTES >> 4));
            if (tc::elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k)                   // UMMA K = 8: 8 TMEM columns of A, 32 bytes of each B row
                    tc::umma_tf32_ts(tmem_big, a_hi + k * 8, d_bhi + soff + uint64_t(k * 2), idesc, (cpos | k) != 0);
                tc::umma_commit(&empty[s]);
                tc::umma_commit(&a_free[it & 1]);
                if (cpos == chunk_iters - 1 || it == iters - 1) tc::umma_commit(&chunk_full[c & 1]);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ++round; }
        }
    } else if (warp == 10) {
        // ================= MMA issuer, small terms: a_lo*b_hi + a_hi*b_lo into the tile-long accumulator =================
        constexpr uint32_t idesc = tc::umma_idesc_tf32(TC_BM, BN);
        const uint32_t stage0 = tc::smem_u32(smem);
        const uint64_t d_bhi = tc::umma_desc_kmajor_sw128(stage0 + TC_A_BYTES, 1024);
        const uint64_t d_blo = tc::umma_desc_kmajor_sw128(stage0 + TC_A_BYTES + Cfg::B_BYTES, 1024);
        int s = 0, round = 0;
        for (int it = 0; it < iters; ++it) {
            const uint32_t a_hi = tmem_a + uint32_t(it & 1) * 64, a_lo = a_hi + 32;
            tc::mbar_wait(&ready[s], round & 1);
            tc::tcgen05_fence_after();
            const uint64_t soff = uint64_t(uint32_t(s) * uint32_t(Cfg::STAGE_BYTES >> 4));
            if (tc::elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    const uint64_t koff = soff + uint64_t(k * 2);
                    tc::umma_tf32_ts(tmem_small, a_lo + k * 8, d_bhi + koff, idesc, (it | k) != 0);
                    tc::umma_tf32_ts(tmem_small, a_hi + k * 8, d_blo + koff, idesc, 1);
                }
                tc::umma_commit(&empty[s]);
                tc::umma_commit(&a_free[it & 1]);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ++round; }
        }
        if (iters > 0 && tc::elect_one_sync()) tc::umma_commit(small_full);
        __syncwarp();
    } else if (warp < 6) {
        // ================= operand transform =================
        // Thread (warp w, lane l) owns row m = 32 (w & 3) + l of the A tile (= TMEM lane m).  It reads the row's 32 floats
        // from the landed tile (undoing the 128-byte swizzle), splits them into TF32 hi / lo and stores both into tensor
        // memory: the MMA then takes A from TMEM, which removes the A_lo round trip and all A operand reads from shared
        // memory -- the kernel was shared-memory-bandwidth bound (measured: 192 KB of smem traffic per stage at 128 B/clk).
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const uint32_t lane_base = uint32_t(q * 32) << 16;
        int s = 0, round = 0;
        for (int it = 0; it < iters; ++it) {
            tc::mbar_wait(&full[s], round & 1);
            if (dbg_flags & 2) {
                tc::mbar_wait(&a_free[it & 1], (((it >> 1) & 1) ^ 1));
                tc::mbar_arrive(&ready[s]);
                if (++s == STAGES) { s = 0; ++round; }
                continue;
            }
            const uint8_t* arow = smem + s * Cfg::STAGE_BYTES + m * 128;
            float wgt = 1.0f;
            if (MODE == MODE_STYLE) wgt = cls_w[nth_set_bit(active, it / kchunks) * TC_BM + m];
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (m & 7)) << 4));
                if (MODE == MODE_STYLE) { v.x *= wgt; v.y *= wgt; v.z *= wgt; v.w *= wgt; }
                const float h0 = tc::round_tf32(v.x), h1 = tc::round_tf32(v.y), h2 = tc::round_tf32(v.z), h3 = tc::round_tf32(v.w);
                hi[c * 4 + 0] = __float_as_uint(h0); lo[c * 4 + 0] = __float_as_uint(tc::round_tf32(v.x - h0));
                hi[c * 4 + 1] = __float_as_uint(h1); lo[c * 4 + 1] = __float_as_uint(tc::round_tf32(v.y - h1));
                hi[c * 4 + 2] = __float_as_uint(h2); lo[c * 4 + 2] = __float_as_uint(tc::round_tf32(v.z - h2));
                hi[c * 4 + 3] = __float_as_uint(h3); lo[c * 4 + 3] = __float_as_uint(tc::round_tf32(v.w - h3));
            }
            // the MMAs that read this TMEM slot two iterations ago must have retired
            tc::mbar_wait(&a_free[it & 1], (((it >> 1) & 1) ^ 1));
            tc::tcgen05_fence_after();
            const uint32_t dst = tmem_a + uint32_t(it & 1) * 64 + lane_base;
            tc::tmem_st_32x32(dst, hi);
            tc::tmem_st_32x32(dst + 32, lo);
            tc::tmem_st_wait();
            tc::tcgen05_fence_before();
            tc::mbar_arrive(&ready[s]);
            if (++s == STAGES) { s = 0; ++round; }
        }
    } else if (warp < 10) {
        // ================= drain (chunk promotion) + epilogue =================
        const int q = warp & 3;                                        // TMEM lane quarter this warp may read
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
                if (dbg_flags & 1) break;
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
        const int m = q * 32 + lane;                                   // accumulator row = pixel within the tile
        const int gy = y0 + m / TC_TW, gx = x0 + m % TC_TW;
        const bool inb = gy < H && gx < W;
        const size_t rowoff = (size_t(gy) * W + gx) * size_t(Cout) + n0;
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
            if (inb) {
                float r[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = acc[c0 + j] + __uint_as_float(v[j]);
                if (MODE == MODE_STYLE) {
                    if (seed != nullptr) {                                 // accumulate into an existing gradient seed
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 sd = *reinterpret_cast<const float4*>(seed + rowoff + c0 + j);
                            r[j] += sd.x; r[j + 1] += sd.y; r[j + 2] += sd.z; r[j + 3] += sd.w;
                        }
                    }
                } else if (MODE == MODE_FWD) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + j));
                        r[j] = fmaxf(r[j] + b.x, 0.f); r[j + 1] = fmaxf(r[j + 1] + b.y, 0.f);
                        r[j + 2] = fmaxf(r[j + 2] + b.z, 0.f); r[j + 3] = fmaxf(r[j + 3] + b.w, 0.f);
                    }
                } else {
                    if (seed != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 sd = __ldg(reinterpret_cast<const float4*>(seed + rowoff + c0 + j));
                            r[j] += sd.x; r[j + 1] += sd.y; r[j + 2] += sd.z; r[j + 3] += sd.w;
                        }
                    }
                    if (mask_src != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 mk = __ldg(reinterpret_cast<const float4*>(mask_src + rowoff + c0 + j));
                            r[j] = mk.x > 0.f ? r[j] : 0.f; r[j + 1] = mk.y > 0.f ? r[j + 1] : 0.f;
                            r[j + 2] = mk.z > 0.f ? r[j + 2] : 0.f; r[j + 3] = mk.w > 0.f ? r[j + 3] : 0.f;
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(Y + rowoff + c0 + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
            }
        }
        tc::tcgen05_fence_before();
    }
    __syncthreads();
    if (trace && threadIdx.x == 0) dbg[4 * 4096 + 1] = clock64();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight preparation: K-major [tap][N][K] hi / lo split
//   forward : N = Cout, K = Cin :  B[tap][co][ci] = W[tap][ci][co]
//   gradient: N = Cin,  K = Cout:  B[tap][ci][co] = W[8-tap][ci][co]
// ---------------------------------------------------------------------------------------------------------------
__global__ void split_weights_kernel(const float* __restrict__ Wf, float* __restrict__ hi, float* __restrict__ lo, int Cin,
                                     int Cout, int gradient) {
    const size_t total = size_t(9) * Cin * Cout;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        float w;
        if (!gradient) {
            const int ci = int(i % Cin);
            const size_t r = i / Cin;
            const int co = int(r % Cout), tap = int(r / Cout);
            w = Wf[(size_t(tap) * Cin + ci) * Cout + co];
        } else {
            const int co = int(i % Cout);
            const size_t r = i / Cout;
            const int ci = int(r % Cin), tap = int(r / Cin);
            w = Wf[(size_t(8 - tap) * Cin + ci) * Cout + co];
        }
        const float h = tc::round_tf32(w);
        hi[i] = h;
        lo[i] = tc::round_tf32(w - h);
    }
}

int prepare_tc_weights(adpst_vgg* h, int i, cudaStream_t st) {
    const int cin = conv_cin(i), cout = conv_cout(i);
    const size_t n = size_t(9) * cin * cout;
    for (int g = 0; g < 2; ++g) {
        ADPST_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->tc_hi[g][i]), n * 4));
        ADPST_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->tc_lo[g][i]), n * 4));
        split_weights_kernel<<<unsigned((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048), 256, 0, st>>>(
            h->wf[i], h->tc_hi[g][i], h->tc_lo[g][i], cin, cout, g);
        ADPST_LAUNCH_CHECK();
        // tensor map of the [9*N][K] matrix (K innermost)
        const int N = g ? cin : cout, K = g ? cout : cin;
        const int BN = N >= 128 ? 128 : N;
        const uint64_t dims[2] = {uint64_t(K), uint64_t(9) * N};
        const uint64_t strides[1] = {uint64_t(K) * 4};
        const uint32_t box[2] = {uint32_t(TC_BK), uint32_t(BN)};
        int rc = tc::make_tensor_map_f32(&h->tm_hi[g][i], h->tc_hi[g][i], 2, dims, strides, box);
        if (rc != ADPST_OK) return rc;
        rc = tc::make_tensor_map_f32(&h->tm_lo[g][i], h->tc_lo[g][i], 2, dims, strides, box);
        if (rc != ADPST_OK) return rc;
    }
    return ADPST_OK;
}

bool conv_tc_eligible(int Cin, int Cout) { return Cin % TC_BK == 0 && (Cout == 64 || Cout % 128 == 0); }

static long long* g_trace_buf = nullptr;
static int g_trace_block = -1;
// timing experiments only (results are wrong when set): 1 = drain skips tcgen05.ld, 2 = transform skips its work,
// 4 = producer skips the B loads, 8 = producer skips the A load
static int g_dbg_flags = getenv("ADPST_TC_DBG") ? atoi(getenv("ADPST_TC_DBG")) : 0;
void conv_tc_set_trace(long long* buf, int block) { g_trace_buf = buf; g_trace_block = block; }

template <int BN, int MODE>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, const float* bias, float* Y,
                     const float* seed, const float* mask, int H, int W, int Cin, int Cout, cudaStream_t st,
                     const float* cls_masks = nullptr, int num_cls = 0) {
    using Cfg = TcCfg<BN>;
    auto kern = conv3x3_tc_kernel<BN, MODE>;
    static bool configured = false;
    if (!configured) {
        ADPST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int tw = (W + TC_TW - 1) / TC_TW, th = (H + TC_TH - 1) / TC_TH;
    dim3 grid(tw * th, Cout / BN);
    kern<<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmBhi, tmBlo, bias, Y, seed, mask, H, W, Cin, Cout, tw, cls_masks,
                                                    num_cls, g_trace_buf, g_trace_block, g_dbg_flags);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

// X: (H,W,Cin) activation; gradient = 0: conv i forward (bias + ReLU), 1: data gradient of conv i (Cin/Cout are the GEMM's
// K and N, i.e. already swapped for the gradient).
int launch_conv_tc(adpst_vgg* h, int i, int gradient, const float* X, float* Y, const float* seed, const float* mask, int H,
                   int W, int Cin, int Cout, cudaStream_t st) {
    CUtensorMap tmA;
    const uint64_t dims[4] = {uint64_t(Cin), uint64_t(W), uint64_t(H), 1};
    const uint64_t strides[3] = {uint64_t(Cin) * 4, uint64_t(W) * Cin * 4, uint64_t(H) * W * Cin * 4};
    const uint32_t box[4] = {uint32_t(TC_BK), uint32_t(TC_TW), uint32_t(TC_TH), 1};
    int rc = tc::make_tensor_map_f32(&tmA, X, 4, dims, strides, box);
    if (rc != ADPST_OK) return rc;
    const CUtensorMap& bh = h->tm_hi[gradient][i];
    const CUtensorMap& bl = h->tm_lo[gradient][i];
    const float* bias = gradient ? nullptr : h->bias[i];
    const int BN = Cout >= 128 ? 128 : Cout;
    if (!gradient) {
        if (BN == 128) return launch_tc<128, MODE_FWD>(tmA, bh, bl, bias, Y, seed, mask, H, W, Cin, Cout, st);
        return launch_tc<64, MODE_FWD>(tmA, bh, bl, bias, Y, seed, mask, H, W, Cin, Cout, st);
    }
    if (BN == 128) return launch_tc<128, MODE_BWD>(tmA, bh, bl, bias, Y, seed, mask, H, W, Cin, Cout, st);
    return launch_tc<64, MODE_BWD>(tmA, bh, bl, bias, Y, seed, mask, H, W, Cin, Cout, st);
}

// ---------------------------------------------------------------------------------------------------------------
// style gradient (components/loss.py:104-137, backward):  dF[px,:] (=|+=) sum_k m_k[px]^2 F[px,:] D_k
// The same kernel with the classes in the role of the filter taps: A = F tile scaled per pixel by m_k^2 in the transform
// warps, B = D_k (symmetric, so K-major == MN-major).  Classes whose mask vanishes on the whole 8x16 pixel tile are
// skipped (exact).
// ---------------------------------------------------------------------------------------------------------------
bool style_tc_eligible(int C) { return C == 64 || C % 128 == 0; }

int launch_style_dF_tc(const float* F, int H, int W, int C, const float* masks, int K, const float* D_hi, const float* D_lo,
                       float* dF, int accumulate, cudaStream_t st) {
    ADPST_REQUIRE(K >= 1 && K <= TC_MAX_CLASSES, "style gradient: K=%d classes not supported (max %d)", K, TC_MAX_CLASSES);
    CUtensorMap tmA, tmH, tmL;
    const uint64_t dims[4] = {uint64_t(C), uint64_t(W), uint64_t(H), 1};
    const uint64_t strides[3] = {uint64_t(C) * 4, uint64_t(W) * C * 4, uint64_t(H) * W * C * 4};
    const uint32_t box[4] = {uint32_t(TC_BK), uint32_t(TC_TW), uint32_t(TC_TH), 1};
    int rc = tc::make_tensor_map_f32(&tmA, F, 4, dims, strides, box);
    if (rc != ADPST_OK) return rc;
    const int BN = C >= 128 ? 128 : C;
    const uint64_t ddims[2] = {uint64_t(C), uint64_t(K) * C};
    const uint64_t dstr[1] = {uint64_t(C) * 4};
    const uint32_t dbox[2] = {uint32_t(TC_BK), uint32_t(BN)};
    rc = tc::make_tensor_map_f32(&tmH, D_hi, 2, ddims, dstr, dbox);
    if (rc != ADPST_OK) return rc;
    rc = tc::make_tensor_map_f32(&tmL, D_lo, 2, ddims, dstr, dbox);
    if (rc != ADPST_OK) return rc;
    const float* seed = accumulate ? dF : nullptr;
    if (BN == 128) return launch_tc<128, MODE_STYLE>(tmA, tmH, tmL, nullptr, dF, seed, nullptr, H, W, C, C, st, masks, K);
    return launch_tc<64, MODE_STYLE>(tmA, tmH, tmL, nullptr, dF, seed, nullptr, H, W, C, C, st, masks, K);
}

}  // namespace adpst
