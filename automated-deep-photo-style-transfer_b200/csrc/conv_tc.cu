// conv_tc.cu -- 3x3 SAME convolution (forward and data gradient) as an implicit GEMM on the 5th-generation tensor
// cores: TMA -> shared memory -> tcgen05.mma (kind::f16, three FP16 terms per product) -> TMEM -> fused epilogue -> HBM.
//
// Replaces the Keras Conv2D calls behind components/VGG19/model.py:30 and their tape gradient (style_transfer.py:341).
//
// GEMM view per CTA:  D[128 pixels, BN channels] = sum over (tap, Cin chunk of 64)  A[128, 64] * B[BN, 64]^T
//   A: an 8 x 16 pixel tile of the NHWC float32 activation, shifted by the tap; two 4-D TMA boxes (C=32, W=16, H=8, N=1)
//      land it with the 128-byte swizzle, and TMA's out-of-bounds zero fill IS the SAME padding, so there is no im2col
//      buffer and no border code.
//   B: weights pre-arranged [tap][Cout][Cin] (K-major) as FP16 hi / lo planes, 2-D TMA box (64, BN), 128-byte swizzle.
//
// Precision: the reference computes in float32 and parity is 1e-5.  One TF32 or FP16 MMA (11 significant
// bits) is 1e-3.  Every product is therefore formed from three tensor-core terms  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi
// accumulated in fp32.  The terms run as kind::f16 (FP16 operands), which the tensor core executes at TWICE the TF32
// rate: FP16 has the same 11 significant bits as TF32, and its narrow exponent is handled by a power-of-two scale per
// tensor derived from the tensor's largest magnitude (tc_common.cuh, "FP16 operand splitting"):
//   t = a * s_a (exact),  a_hi = fp16(t),  a_lo = fp16((t - a_hi) * 2^11);  the same for the weights with s_b.
//   big   = sum a_hi*b_hi                      (scaled by s_a s_b)
//   small = sum a_lo*b_hi + a_hi*b_lo           (scaled by s_a s_b 2^11)
//   out   = (big + small * 2^-11) / (s_a s_b)   -- all rescales are powers of two, hence exact.
// Every producer of an activation or gradient tensor records max|.| with one atomicMax per warp (the conv epilogue
// here, the pool / ReLU-mask kernels in vgg_simt.cu), so the next layer finds its scale in device memory: no host
// round trip, CUDA-graph replayable.
//   b_hi / b_lo are split once when the weights are loaded; a_hi / a_lo are split by the "transform" warps between the
//   TMA landing and the MMA issue and handed to the tensor core through TENSOR MEMORY (tcgen05.st; the MMA's A operand
//   is a TMEM address), not shared memory.
// The dropped a_lo*b_lo term and the rounding of the lo parts are O(2^-22) relative.
//
// Warp roles (352 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer for the big term, warp 10 = MMA
// issuer for the small terms, warps 2..5 = operand transform (A hi/lo split into tensor memory), warps 6..9 = chunk
// promotion TMEM -> registers, then the epilogue (bias/ReLU or seed/mask -> global).  mbarrier rings: full (TMA landed) -> ready (A split done) -> empty (MMAs retired),
// chunk_full / chunk_empty for the two TMEM chunk buffers.
#include <cuda_fp16.h>
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

static int make_tensor_map(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(ADPST_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t d[5], s[5];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = fn(out, dt, cuuint32_t(rank), const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ADPST_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return ADPST_OK;
}
int make_tensor_map_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box) {
    return make_tensor_map(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}
int make_tensor_map_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box) {
    return make_tensor_map(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, rank, dims, strides_bytes, box);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------------------------
// Accumulation accuracy.  Measured on B200 (scripts/tc_error_probe.py): tcgen05.mma adds into its fp32 accumulator with
// truncation, a systematic bias of about -2^-26 of the accumulator per MMA; over a whole K loop (hundreds of MMAs) that
// is above the 1e-5 parity bar.  Therefore:
//   * the big term a_hi*b_hi is accumulated in TMEM only over CHUNK_ITERS stages (8 MMAs) at a time, in two TMEM buffers
//     used alternately; four "drain" warps promote each finished chunk into fp32 REGISTERS with round-to-nearest adds
//     while the tensor core already works on the next chunk;
//   * the two small terms (2^-11 of the big one) accumulate in a third TMEM buffer for the whole tile -- their bias is
//     2^-11 smaller and negligible -- and are added once at the end.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TC_TH = 8, TC_TW = 16, TC_BM = TC_TH * TC_TW;     // 128 pixels = UMMA M
constexpr int TC_BK = 64;                                        // K per stage: 64 fp16 = one 128-byte swizzle row of B
constexpr int TC_THREADS = 512;                                  // 4 warpgroups: {TMA, MMA big, MMA small, -}, transform x2, drain
// registers per thread after setmaxnreg (the launch gives every thread 65536 / 512 = 128):
constexpr int TC_REGS_CTRL = 40, TC_REGS_XFORM = 112, TC_REGS_DRAIN = 224;     // 128 * (40 + 2 * 112 + 224) = 62464 <= 65536
constexpr int TC_HW = TC_TW + 2, TC_HH = TC_TH + 2;              // the tile plus its 1-pixel halo: 18 x 10 pixels
constexpr int TC_A_BOX_BYTES = 23 * 1024;                        // one landed box: 180 halo pixels x 32 channels fp32 = 23040 B,
                                                                 // padded to the 1024-byte alignment of the 128-byte swizzle
constexpr int TC_A_TX_BYTES = 2 * TC_HW * TC_HH * 32 * 4;        // bytes TMA actually delivers per A stage
constexpr int TC_A_BYTES = 2 * TC_A_BOX_BYTES;                   // 46 KB: channels [0,32) and [32,64) of the K chunk
constexpr int TC_CHUNK_ITERS = 2;                                // stages per promoted chunk (8 big MMAs)
constexpr int TC_STYLE_PRELOAD = 4;                              // style mode: class weights preloaded per work item
constexpr int TC_MAX_CLASSES = 32;                               // style mode: classes per launch (active set is a bit mask)

// Shared-memory bandwidth (128 B/clk/SM) is what bounds this kernel, not the tensor pipe: per K chunk of 64 and BN = 128
// the MMAs read 48 KB of B (one operand per MMA, 12 MMAs) in 768 clk, TMA writes 32 KB of B, and the A operand costs its
// TMA write plus the transform warps' reads.  So the A tile is NOT reloaded per filter tap: one (8+2) x (16+2) pixel halo
// tile per K chunk serves all nine taps (the transform warps read their row at the tap's offset), which cuts the A bytes
// written to shared memory -- and fetched from L2 -- from 9 x 32 KB to 45 KB per chunk.
// Two rings: the A halo tile lives for nine stages, the B tiles until the MMAs that use them retire.
template <int BN, int MODE> struct TcCfg {
    // Ring depths.  Measured alternatives that did NOT pay (round 2, 1024^2 step): three halo-tile slots with a 4-deep B ring
    // for BN = 64 (block1_conv2 269/279 TFLOP/s instead of 280/285), two B slots + three halo-tile slots for the style
    // gradient (330 us instead of 312 us over the four BN = 128 launches), a second small-term accumulator for BN = 64.
    static constexpr int A_STAGES = 2;
    static constexpr int B_STAGES = BN == 128 ? (MODE == MODE_BWD ? 4 : 3) : 6;
    static constexpr int B_BYTES = BN * TC_BK * 2;                          // BN rows x 64 fp16
    static constexpr int B_STAGE_BYTES = 2 * B_BYTES;                       // B_hi, B_lo
    static constexpr int OFF_B = A_STAGES * TC_A_BYTES;
    static constexpr int OFF_BARS = OFF_B + B_STAGES * B_STAGE_BYTES;
    // The drain warps transpose their 32 pixel x 32 channel pieces through shared memory so that a store instruction writes four
    // complete 128-byte lines instead of 32 half sectors a pixel apart (the L1 data pipe, which also carries the transform
    // warps' shared-memory loads, is what the layers with few K chunks wait for).  BN = 128 pays for the staging tile with one B
    // stage (3 instead of 4).  Forward and style-gradient kernels only (one train step at 1024^2, ncu: forward launches
    // 1708 + 333 -> 1669 + 313 us, style gradient 318 + 154 -> 306 + 134 us).  The data-gradient kernels keep the direct stores:
    // their epilogue also loads the ReLU mask and the seed, the __syncwarp()s of the staging tile stop the compiler from issuing
    // those loads ahead of the chunk loop, and the exposed latency cost more than the stores saved (1764 + 584 -> 1868 + 620 us;
    // routing the loads through the tile as well: 2031 + 659 us).
    static constexpr bool STAGE_OUT = (MODE != MODE_BWD);
    static constexpr int OFF_STAGE = OFF_BARS + 256;
    static constexpr int STAGE_BYTES = STAGE_OUT ? 4 * 32 * 33 * 4 : 0;
    static constexpr int SMEM_BYTES = OFF_BARS + 1024 /*alignment slack*/ + 256 /*barriers*/ + STAGE_BYTES;
    // tensor memory columns: big0 | big1 | small | A operand slots (each: hi = 32 columns of packed fp16 pairs, lo = 32).
    // BN = 128 fills the 512 columns with two slots; BN = 64 has room for four, which it needs: its MMAs of one stage take
    // 384 clk, less than the round trip  tcgen05.st -> MMA -> commit -> a_free -> next tcgen05.st  of a two-slot ring.
    static constexpr int A_SLOTS = BN == 128 ? 2 : 4;
    static constexpr int SMALL_BUFS = 1;                                    // small-term accumulators (2 fit for BN = 64: no gain)
    static constexpr uint32_t COL_SMALL = 2 * BN, COL_A = (2 + SMALL_BUFS) * BN, TMEM_COLS = 512;
    static_assert(COL_A + A_SLOTS * 64 <= TMEM_COLS, "tensor memory columns");
};

// position of the n-th (0-based) set bit of m
__device__ __forceinline__ int nth_set_bit(uint32_t m, int n) {
    for (int i = 0; i < n; ++i) m &= m - 1;
    return __ffs(int(m)) - 1;
}

__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// Persistent kernel: gridDim.x CTAs (one per SM) walk the work items (pixel tile, block of BN output channels) with
// stride gridDim.x; the output-channel block is the fastest index, so CTAs that run side by side read the same A tile
// (L2 hits) and all of them the same weights.  Every warp role keeps its pipeline position (stage, phase, chunk buffer)
// across work items, so the TMA loads, operand splits and MMAs of the next item run while the drain warps are still
// writing the previous item's outputs: per-item prologue/epilogue cost is hidden, which is what the short K loops
// (Cin = 64: nine stages per item; style gradient: one or two) need.
template <int BN, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                  const __grid_constant__ CUtensorMap tmBlo, const float* __restrict__ bias, float* __restrict__ Y,
                  const float* __restrict__ seed, const float* __restrict__ mask_src, int H, int W, int Cin, int Cout,
                  int tiles_w, int num_tiles, const float* __restrict__ cls_masks, const uint32_t* __restrict__ tile_active,
                  const uint32_t* __restrict__ a_absmax, const uint32_t* __restrict__ b_absmax,
                  const uint32_t* __restrict__ w_absmax, uint32_t* __restrict__ y_absmax, float* __restrict__ pool_out,
                  int tw_log2_arg, int pool_pitch, int pool_xoff) {
    using Cfg = TcCfg<BN, MODE>;
    // pixel tile: 8 rows x 16 columns (tw_log2 = 4) or 16 rows x 8 columns (tw_log2 = 3, narrow maps: the column strips of a
    // spatially tiled run are 68 or 34 pixels wide at 1/8 and 1/16 resolution); 128 pixels and a 180-pixel halo tile either way
    const int tw_log2 = (MODE == MODE_STYLE) ? 4 : tw_log2_arg;             // (the per-tile class sets are made for 8 x 16)
    constexpr int AST = Cfg::A_STAGES, BST = Cfg::B_STAGES;
    constexpr int SB = Cfg::SMALL_BUFS;         // small-term accumulators in tensor memory
    constexpr int NSLOT = Cfg::A_SLOTS;         // A operand slots in tensor memory (ring between the transform warps and the MMAs)
    static_assert(BN == 64 || BN == 128, "tile width");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
    uint64_t* full = bars;                      // [AST]  A tile landed (the transform warps start on it at once)
    uint64_t* a_empty = bars + 3;               // [AST]  the transform warps have read the A tile
    uint64_t* ready = bars + 6;                 // [BST]  B tiles landed and A split into hi/lo in tensor memory
    uint64_t* empty = bars + 12;                // [BST]  MMAs that read the B tiles have retired
    uint64_t* chunk_full = bars + 18;           // [2]    a big-term chunk is complete in TMEM buffer b
    uint64_t* chunk_empty = chunk_full + 2;     // [2]       the drain warps have consumed TMEM buffer b
    uint64_t* small_full = chunk_empty + 2;     // [SB]      every MMA of the work item has retired
    uint64_t* small_empty = small_full + 2;     // [SB]      the drain warps have read the small-term accumulator
    uint64_t* a_free = small_empty + 2;         // [NSLOT]   the MMAs that read TMEM A slot j have retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = Cin / TC_BK;
    const int nblk = Cout / BN;
    const int total = num_tiles * nblk;
    constexpr int chunk_iters = TC_CHUNK_ITERS;
    // operand scales (powers of two): s_a from the input tensor's recorded max|.|, s_b from the weights'
    int ea = tc::f16_scale_exponent(__ldg(a_absmax));
    const int eb = tc::f16_scale_exponent(__ldg(b_absmax));
    if (MODE == MODE_STYLE) {
        const uint32_t wb = __ldg(w_absmax);                           // class weights above 1 (masks outside [0,1])
        if (wb > 0x3F800000u) ea -= int(wb >> 23) - 127 + 1;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < AST; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&a_empty[s], 256);        // both transform warpgroups
        }
        for (int s = 0; s < BST; ++s) {
            tc::mbar_init(&ready[s], 128 + 1);      // 128 transform threads + the producer's expect_tx arrival (B bytes)
            tc::mbar_init(&empty[s], 2);            // both MMA-issuing warps commit
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&chunk_full[b], 1);
            tc::mbar_init(&chunk_empty[b], 128);
        }
        for (int j = 0; j < SB; ++j) {
            tc::mbar_init(&small_full[j], 1);
            tc::mbar_init(&small_empty[j], 128);
        }
        for (int j = 0; j < NSLOT; ++j) tc::mbar_init(&a_free[j], 2);
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
    const uint32_t tmem_small0 = tmem_base + Cfg::COL_SMALL;
    const uint32_t tmem_a = tmem_base + Cfg::COL_A;

    // "taps" of an item's K loop: the 9 filter taps (convolution) or the classes whose mask is non-zero somewhere in the
    // pixel tile (style gradient: dF = sum_k w_k F D_k is a 1x1 convolution per class with a per-pixel weight
    // w_k = m_k^2; the per-tile class sets come from style_tiles_kernel).
    auto item_active = [&](int tile) -> uint32_t { return MODE == MODE_STYLE ? __ldg(tile_active + tile) : 0x1FFu; };

    // Register budget per role (setmaxnreg, whole warpgroups): the control warps need almost none, the drain warps hold a
    // 128 x BN fp32 tile row.
    if (warp == 0) {
        // ================= TMA producer =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_CTRL));
        if (lane == 0) {
            int sa = 0, ra = 0, s = 0, round = 0;                      // A ring slot / phase, B ring slot / phase
            // The halo tile of a K chunk comes from HBM (1-2 us), its nine B tiles from L2.  The tile of chunk c + AST - 1 is
            // therefore requested while the B tiles of chunk c are still being issued -- after the first BST of them, by which
            // time the MMAs of chunk c - 1 have retired and its slot is free without waiting -- instead of after the last one.
            // With one K chunk per work item (Cin = 64) the previous order (request after the last B tile, two slots) exposed
            // most of the load latency once per item (ncu: tensor pipe 40 % busy on block1_conv2).
            int wa = blockIdx.x, kca = 0;                              // cursor of the next halo tile to request
            auto skip_empty = [&]() { while (wa < total && __popc(item_active(wa / nblk)) == 0) wa += gridDim.x; };
            auto issue_a = [&]() {
                if (wa >= total) return;
                const int tile = wa / nblk;
                const int y0 = (tile / tiles_w) << (7 - tw_log2), x0 = (tile % tiles_w) << tw_log2;
                // 10 x 18 pixels x 64 channels, zero-filled outside the image (= SAME padding)
                tc::mbar_wait(&a_empty[sa], (ra & 1) ^ 1);
                uint8_t* sta = smem + sa * TC_A_BYTES;
                tc::mbar_arrive_expect_tx(&full[sa], TC_A_TX_BYTES);
                tc::tma_load_4d(sta, &tmA, &full[sa], kca * TC_BK, x0 - 1, y0 - 1, 0);
                tc::tma_load_4d(sta + TC_A_BOX_BYTES, &tmA, &full[sa], kca * TC_BK + 32, x0 - 1, y0 - 1, 0);
                if (++sa == AST) { sa = 0; ++ra; }
                if (++kca == kchunks) { kca = 0; wa += gridDim.x; skip_empty(); }
            };
            skip_empty();
            for (int j = 0; j < AST - 1; ++j) issue_a();              // run AST - 1 halo tiles ahead of the B tiles
            for (int w = blockIdx.x; w < total; w += gridDim.x) {
                const int tile = w / nblk, n0 = (w - tile * nblk) * BN;
                const uint32_t active = item_active(tile);
                const int ntaps = __popc(active);
                if (ntaps == 0) continue;
                const int a_at = (ntaps < BST ? ntaps : BST) - 1;      // request the next halo tile after this many B tiles
                for (int kc = 0; kc < kchunks; ++kc) {
                    for (int slot = 0; slot < ntaps; ++slot) {
                        const int tap = (MODE == MODE_STYLE) ? nth_set_bit(active, slot) : slot;
                        tc::mbar_wait(&empty[s], (round & 1) ^ 1);
                        uint8_t* stb = smem + Cfg::OFF_B + s * Cfg::B_STAGE_BYTES;
                        tc::mbar_arrive_expect_tx(&ready[s], 2 * Cfg::B_BYTES);
                        tc::tma_load_2d(stb, &tmBhi, &ready[s], kc * TC_BK, tap * Cout + n0);
                        tc::tma_load_2d(stb + Cfg::B_BYTES, &tmBlo, &ready[s], kc * TC_BK, tap * Cout + n0);
                        if (++s == BST) { s = 0; ++round; }
                        if (slot == a_at) issue_a();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer, big term =================
        // Issuing one tcgen05.mma costs the issuing thread 40-48 clk (measured, tests/cuda/umma_queue_probe.cu) while a
        // 128x128x16 FP16 MMA executes in 64 clk, so ONE thread that also has ~240 clk of barrier work per stage cannot
        // keep the pipe fed with 12 MMAs per stage.  The issue is therefore split over two warps that own DIFFERENT
        // accumulators (no cross-thread ordering needed): this warp issues a_hi*b_hi into the chunk buffers, warp 2
        // issues the two small terms.  Both commit to empty[] / a_free[] (arrival count 2).
        // The whole warp runs the loop (warp-uniform control flow, descriptors in uniform registers); one elected lane
        // issues.  Descriptors are built once: per stage and K-step only the 14-bit address field changes.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_CTRL));
        constexpr uint32_t idesc = tc::umma_idesc_f16(TC_BM, BN);
        const uint32_t stage0 = tc::smem_u32(smem);
        const uint64_t d_bhi = tc::umma_desc_kmajor_sw128(stage0 + Cfg::OFF_B, 1024);
        int s = 0, round = 0, git = 0, gc = 0;                         // stage / phase, global iteration, global chunk
        for (int w = blockIdx.x; w < total; w += gridDim.x) {
            const int iters = __popc(item_active(w / nblk)) * kchunks;
            for (int it = 0; it < iters; ++it, ++git) {
                const int cpos = it % chunk_iters;
                const uint32_t tmem_big = tmem_base + uint32_t(gc & 1) * BN;
                const uint32_t a_hi = tmem_a + uint32_t(git % NSLOT) * 64;
                if (cpos == 0) tc::mbar_wait(&chunk_empty[gc & 1], ((gc >> 1) & 1) ^ 1);   // TMEM buffer drained
                // ONE wait per stage: ready[s] completes when the B tiles have landed (transaction bytes) and the 128
                // transform threads have stored A hi/lo into tensor memory.
                tc::mbar_wait(&ready[s], round & 1);
                tc::tcgen05_fence_after();
                const uint64_t soff = uint64_t(uint32_t(s) * uint32_t(Cfg::B_STAGE_BYTES >> 4));
                const bool close = (cpos == chunk_iters - 1 || it == iters - 1);
                if (tc::elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)              // UMMA K = 16: 8 TMEM columns of A, 32 bytes of each B row
                        tc::umma_f16_ts(tmem_big, a_hi + k * 8, d_bhi + soff + uint64_t(k * 2), idesc, (cpos | k) != 0);
                    tc::umma_commit(&empty[s]);
                    tc::umma_commit(&a_free[git % NSLOT]);
                    if (close) tc::umma_commit(&chunk_full[gc & 1]);
                }
                __syncwarp();
                if (close) ++gc;
                if (++s == BST) { s = 0; ++round; }
            }
        }
    } else if (warp == 2) {
        // ================= MMA issuer, small terms: a_lo*b_hi + a_hi*b_lo into the item-long accumulator =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_CTRL));
        constexpr uint32_t idesc = tc::umma_idesc_f16(TC_BM, BN);
        const uint32_t stage0 = tc::smem_u32(smem);
        const uint64_t d_bhi = tc::umma_desc_kmajor_sw128(stage0 + Cfg::OFF_B, 1024);
        const uint64_t d_blo = tc::umma_desc_kmajor_sw128(stage0 + Cfg::OFF_B + Cfg::B_BYTES, 1024);
        int s = 0, round = 0, git = 0, sj = 0;                         // sj: items with a non-empty K loop so far
        for (int w = blockIdx.x; w < total; w += gridDim.x) {
            const int iters = __popc(item_active(w / nblk)) * kchunks;
            if (iters == 0) continue;
            // the accumulator this item uses must have been read by the drain warps (SB items ago; free the first SB times)
            tc::mbar_wait(&small_empty[sj % SB], ((sj / SB) & 1) ^ 1);
            tc::tcgen05_fence_after();
            const uint32_t tmem_small = tmem_small0 + uint32_t(sj % SB) * BN;
            for (int it = 0; it < iters; ++it, ++git) {
                const uint32_t a_hi = tmem_a + uint32_t(git % NSLOT) * 64, a_lo = a_hi + 32;
                tc::mbar_wait(&ready[s], round & 1);
                tc::tcgen05_fence_after();
                const uint64_t soff = uint64_t(uint32_t(s) * uint32_t(Cfg::B_STAGE_BYTES >> 4));
                if (tc::elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint64_t koff = soff + uint64_t(k * 2);
                        tc::umma_f16_ts(tmem_small, a_lo + k * 8, d_bhi + koff, idesc, (it | k) != 0);
                        tc::umma_f16_ts(tmem_small, a_hi + k * 8, d_blo + koff, idesc, 1);
                    }
                    tc::umma_commit(&empty[s]);
                    tc::umma_commit(&a_free[git % NSLOT]);
                    if (it == iters - 1) tc::umma_commit(&small_full[sj % SB]);
                }
                __syncwarp();
                if (++s == BST) { s = 0; ++round; }
            }
            ++sj;
        }
    } else if (warp == 3) {
        // (idle: completes the control warpgroup so that setmaxnreg applies to whole warpgroups)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_CTRL));
    } else if (warp < 12) {
        // ================= operand transform (two warpgroups, alternating stages) =================
        // Thread (warp w, lane l) owns row m = 32 (w & 3) + l of the A tile (= TMEM lane m).  It reads the row's 64 floats
        // from the halo tile (undoing the 128-byte swizzle), scales them into FP16 range, splits them into hi / lo
        // and stores both, packed two per 32-bit column, into tensor memory: the MMA then takes A from TMEM, which removes
        // every A operand read from shared memory (the kernel was shared-memory-bandwidth bound with A in smem).
        // One stage of this loop (16 loads, ~350 arithmetic instructions, tcgen05.st + wait) takes longer than the 768 clk
        // its MMAs need, so warpgroup g handles the stages with (global stage index & 1) == g and owns TMEM A slot g.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_XFORM));
        constexpr bool PRESPLIT = (MODE != MODE_STYLE);
        const int grp = (warp - 4) >> 2;
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const uint32_t lane_base = uint32_t(q * 32) << 16;
        const float scale_a = tc::pow2f_int(ea);
        const uint32_t smem_base = tc::smem_u32(smem);
        int sa = 0, ra = 0, s = 0, git = 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x) {
            const int tile = w / nblk;
            const uint32_t active = item_active(tile);
            const int ntaps = __popc(active);
            if (ntaps == 0) continue;
            const int halo_w = (1 << tw_log2) + 2;
            const int ty = m >> tw_log2, tx = m & ((1 << tw_log2) - 1);
            const int gy = ((tile / tiles_w) << (7 - tw_log2)) + ty, gx = ((tile % tiles_w) << tw_log2) + tx;
            // style mode: this pixel's weight m_k^2 for the first classes of the tile, requested once per work item (they do
            // not depend on the K chunk) instead of one exposed global load per stage
            float wq[TC_STYLE_PRELOAD];
            if (MODE == MODE_STYLE) {
#pragma unroll
                for (int j = 0; j < TC_STYLE_PRELOAD; ++j) {
                    float wk = 0.f;
                    if (j < ntaps && gy < H && gx < W)
                        wk = cls_masks ? __ldg(cls_masks + size_t(nth_set_bit(active, j)) * H * W + size_t(gy) * W + gx) : 1.0f;
                    wq[j] = wk * wk;
                }
            }
            for (int kc = 0; kc < kchunks; ++kc) {
                tc::mbar_wait(&full[sa], ra & 1);
                if (PRESPLIT) {
                    // Convolution modes: every pixel of the halo tile is used by up to nine taps with the SAME scale, so it is
                    // split into FP16 hi / lo ONCE, in place: row r of box 0 (channels 0..31 as float32, 128 bytes) becomes the
                    // hi halves of all 64 channels, row r of box 1 the lo halves (same 16-byte-chunk swizzle).  A thread reads
                    // both rows completely before it writes, and touches no other row, so the only synchronisation needed is
                    // one named barrier of the 256 transform threads before the taps start reading shifted rows.  The per-tap
                    // work shrinks to a 256-byte copy shared memory -> tensor memory (it was 64 multiplies, 64 conversion
                    // pairs and the packing per tap: the Cout = 64 layers were bound by the issue slots of these warps).
                    const int r = int(threadIdx.x) - 128;                   // transform threads are 128..383
                    if (r < TC_HW * TC_HH) {
                        const uint32_t row0 = smem_base + uint32_t(sa * TC_A_BYTES + r * 128);
                        const uint32_t row1 = row0 + TC_A_BOX_BYTES;
                        float4 v[16];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            v[c] = tc::lds128(row0 + uint32_t((c ^ (r & 7)) << 4));
                            v[8 + c] = tc::lds128(row1 + uint32_t((c ^ (r & 7)) << 4));
                        }
                        uint32_t hi[32], lo[32];
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const float t0 = v[c].x * scale_a, t1 = v[c].y * scale_a, t2 = v[c].z * scale_a, t3 = v[c].w * scale_a;
                            const __half2 h01 = __floats2half2_rn(t0, t1), h23 = __floats2half2_rn(t2, t3);
                            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                            const __half2 l01 = __floats2half2_rn((t0 - f01.x) * 2048.0f, (t1 - f01.y) * 2048.0f);
                            const __half2 l23 = __floats2half2_rn((t2 - f23.x) * 2048.0f, (t3 - f23.y) * 2048.0f);
                            hi[c * 2] = h2_bits(h01); hi[c * 2 + 1] = h2_bits(h23);
                            lo[c * 2] = h2_bits(l01); lo[c * 2 + 1] = h2_bits(l23);
                        }
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            tc::sts128(row0 + uint32_t((c ^ (r & 7)) << 4), hi[c * 4], hi[c * 4 + 1], hi[c * 4 + 2], hi[c * 4 + 3]);
                            tc::sts128(row1 + uint32_t((c ^ (r & 7)) << 4), lo[c * 4], lo[c * 4 + 1], lo[c * 4 + 2], lo[c * 4 + 3]);
                        }
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                for (int slot = 0; slot < ntaps; ++slot, ++git) {
                    if ((git & 1) != grp) {                                 // the other warpgroup's stage
                        if (++s == BST) s = 0;
                        continue;
                    }
                    if (PRESPLIT) {
                        const uint32_t dst = tmem_a + uint32_t(git % NSLOT) * 64 + lane_base;
                        const int kh = slot / 3, kw = slot - kh * 3;
                        const int r = (ty + kh) * halo_w + tx + kw;             // this thread's pixel, shifted by the tap, in the halo tile
                        const uint32_t row0 = smem_base + uint32_t(sa * TC_A_BYTES + r * 128);
                        const uint32_t row1 = row0 + TC_A_BOX_BYTES;
                        uint32_t hi[2][16], lo[2][16];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 h = tc::lds128(row0 + uint32_t((c ^ (r & 7)) << 4));
                            const float4 l = tc::lds128(row1 + uint32_t((c ^ (r & 7)) << 4));
                            hi[c >> 2][(c & 3) * 4] = __float_as_uint(h.x); hi[c >> 2][(c & 3) * 4 + 1] = __float_as_uint(h.y);
                            hi[c >> 2][(c & 3) * 4 + 2] = __float_as_uint(h.z); hi[c >> 2][(c & 3) * 4 + 3] = __float_as_uint(h.w);
                            lo[c >> 2][(c & 3) * 4] = __float_as_uint(l.x); lo[c >> 2][(c & 3) * 4 + 1] = __float_as_uint(l.y);
                            lo[c >> 2][(c & 3) * 4 + 2] = __float_as_uint(l.z); lo[c >> 2][(c & 3) * 4 + 3] = __float_as_uint(l.w);
                        }
                        // the MMAs that read this TMEM slot two iterations ago must have retired
                        tc::mbar_wait(&a_free[git % NSLOT], (((git / NSLOT) & 1) ^ 1));
                        tc::tcgen05_fence_after();
                        tc::tmem_st_32x16(dst, hi[0]);
                        tc::tmem_st_32x16(dst + 16, hi[1]);
                        tc::tmem_st_32x16(dst + 32, lo[0]);
                        tc::tmem_st_32x16(dst + 48, lo[1]);
                        tc::tmem_st_wait();
                        tc::tcgen05_fence_before();
                        tc::mbar_arrive(&ready[s]);
                        if (++s == BST) s = 0;
                        continue;
                    }
                    float scl = scale_a;
                    int kh = 1, kw = 1;
                    if (MODE == MODE_STYLE) {
                        float wk = 0.f;
                        if (slot < TC_STYLE_PRELOAD) {
                            wk = slot == 0 ? wq[0] : slot == 1 ? wq[1] : slot == 2 ? wq[2] : wq[3];
                        } else if (gy < H && gx < W) {
                            wk = cls_masks ? __ldg(cls_masks + size_t(nth_set_bit(active, slot)) * H * W + size_t(gy) * W + gx) : 1.0f;
                            wk *= wk;
                        }
                        scl *= wk;
                    } else {
                        kh = slot / 3; kw = slot - kh * 3;
                    }
                    const int r = (ty + kh) * halo_w + tx + kw;             // this thread's pixel, shifted by the tap, in the halo tile
                    const uint32_t dst = tmem_a + uint32_t(git % NSLOT) * 64 + lane_base;
                    uint32_t hi[2][16], lo[2][16];
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t arow = smem_base + uint32_t(sa * TC_A_BYTES + half * TC_A_BOX_BYTES + r * 128);
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 v = tc::lds128(arow + uint32_t((c ^ (r & 7)) << 4));
                            const float t0 = v.x * scl, t1 = v.y * scl, t2 = v.z * scl, t3 = v.w * scl;
                            const __half2 h01 = __floats2half2_rn(t0, t1), h23 = __floats2half2_rn(t2, t3);
                            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                            const __half2 l01 = __floats2half2_rn((t0 - f01.x) * 2048.0f, (t1 - f01.y) * 2048.0f);
                            const __half2 l23 = __floats2half2_rn((t2 - f23.x) * 2048.0f, (t3 - f23.y) * 2048.0f);
                            hi[half][c * 2] = h2_bits(h01); hi[half][c * 2 + 1] = h2_bits(h23);
                            lo[half][c * 2] = h2_bits(l01); lo[half][c * 2 + 1] = h2_bits(l23);
                        }
                    }
                    // the MMAs that read this TMEM slot two iterations ago must have retired
                    tc::mbar_wait(&a_free[git % NSLOT], (((git / NSLOT) & 1) ^ 1));
                    tc::tcgen05_fence_after();
                    tc::tmem_st_32x16(dst, hi[0]);
                    tc::tmem_st_32x16(dst + 16, hi[1]);
                    tc::tmem_st_32x16(dst + 32, lo[0]);
                    tc::tmem_st_32x16(dst + 48, lo[1]);
                    tc::tmem_st_wait();
                    tc::tcgen05_fence_before();
                    tc::mbar_arrive(&ready[s]);
                    if (++s == BST) s = 0;
                }
                // every tap has consumed the halo tile (the tcgen05.st above needed the loaded values): release its slot
                if (PRESPLIT) tc::fence_proxy_async_smem();                // generic-proxy writes before TMA overwrites the slot
                tc::mbar_arrive(&a_empty[sa]);
                if (++sa == AST) { sa = 0; ++ra; }
            }
        }
    } else {
        // ================= drain (chunk promotion) + epilogue =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_DRAIN));
        const int q = warp & 3;                                        // TMEM lane quarter this warp may read
        const uint32_t lane_base = uint32_t(q * 32) << 16;
        const float inv_big = tc::pow2f_int(-(ea + eb)), inv_small = tc::pow2f_int(-(ea + eb) - 11);
        const int m = q * 32 + lane;                                   // accumulator row = pixel within the tile
        float amax = 0.f;
        int gc = 0, sj = 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x) {
            const int tile = w / nblk, n0 = (w - tile * nblk) * BN;
            const int iters = __popc(item_active(tile)) * kchunks;
            const int nchunks = (iters + chunk_iters - 1) / chunk_iters;
            float acc[BN];
#pragma unroll
            for (int j = 0; j < BN; ++j) acc[j] = 0.f;
            for (int c = 0; c < nchunks; ++c, ++gc) {
                tc::mbar_wait(&chunk_full[gc & 1], (gc >> 1) & 1);
                tc::tcgen05_fence_after();
                const uint32_t src = tmem_base + uint32_t(gc & 1) * BN + lane_base;
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tc::tmem_ld_32x32(src + c0, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
                }
                tc::tcgen05_fence_before();
                tc::mbar_arrive(&chunk_empty[gc & 1]);
            }
            if (iters > 0) {
                // fold in the small terms and hand their accumulator back at once: the next item's MMAs are already running
                const uint32_t tmem_small = tmem_small0 + uint32_t(sj % SB) * BN;
                tc::mbar_wait(&small_full[sj % SB], (sj / SB) & 1);
                tc::tcgen05_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tc::tmem_ld_32x32(tmem_small + lane_base + c0, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + j] = fmaf(__uint_as_float(v[j]), inv_small, acc[c0 + j] * inv_big);
                }
                tc::tcgen05_fence_before();
                tc::mbar_arrive(&small_empty[sj % SB]);
                ++sj;
            }
            const int tile_w = 1 << tw_log2;
            const int gy = ((tile / tiles_w) << (7 - tw_log2)) + (m >> tw_log2), gx = ((tile % tiles_w) << tw_log2) + (m & (tile_w - 1));
            const bool inb = gy < H && gx < W;
            if (Cfg::STAGE_OUT || inb || (MODE == MODE_FWD && pool_out != nullptr)) {   // (pooling shuffles / the staged store need the whole warp)
                const size_t rowoff = (size_t(gy) * W + gx) * size_t(Cout) + n0;
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    float r[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = acc[c0 + j];
                    if (MODE == MODE_STYLE) {
                        if (seed != nullptr && inb) {                      // accumulate into an existing gradient seed (= Y)
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
                        if (pool_out != nullptr) {
                            // fused 2x2/2 VALID max-pool (model.py / Keras MaxPooling2D): the window partners of pixel
                            // (ty, tx) are lanes ^1 (tx + 1) and ^tile_w (ty + 1) of this warp -- a warp holds tile rows 2q, 2q+1
                            // (4q .. 4q+3 of a 16 x 8 tile).  Windows that exist lie completely inside the image, so
                            // out-of-image lanes never contribute.
                            const bool writer = (lane & (tile_w | 1)) == 0 && gy + 1 < H && gx + 1 < W;
                            float* prow = pool_out + (size_t(gy >> 1) * pool_pitch + (gx >> 1) + pool_xoff) * size_t(Cout) + n0 + c0;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float pv[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    float v = fmaxf(r[j + e], __shfl_xor_sync(0xffffffffu, r[j + e], 1));
                                    pv[e] = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, tile_w));
                                }
                                if (writer) *reinterpret_cast<float4*>(prow + j) = make_float4(pv[0], pv[1], pv[2], pv[3]);
                            }
                        }
                    } else {
                        if (seed != nullptr && inb) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 sd = __ldg(reinterpret_cast<const float4*>(seed + rowoff + c0 + j));
                                r[j] += sd.x; r[j + 1] += sd.y; r[j + 2] += sd.z; r[j + 3] += sd.w;
                            }
                        }
                        if (mask_src != nullptr && inb) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 mk = __ldg(reinterpret_cast<const float4*>(mask_src + rowoff + c0 + j));
                                r[j] = mk.x > 0.f ? r[j] : 0.f; r[j + 1] = mk.y > 0.f ? r[j + 1] : 0.f;
                                r[j + 2] = mk.z > 0.f ? r[j + 2] : 0.f; r[j + 3] = mk.w > 0.f ? r[j + 3] : 0.f;
                            }
                        }
                    }
                    if (inb) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) amax = fmaxf(amax, fabsf(r[j]));
                    }
                    if constexpr (Cfg::STAGE_OUT) {
                        // pixel-major registers -> shared memory (row = pixel, 33-float pitch: conflict-free both ways) -> every
                        // store instruction writes 4 pixels x 128 contiguous bytes
                        float* stg = reinterpret_cast<float*>(smem + Cfg::OFF_STAGE) + q * (32 * 33);
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = r[j];
                        __syncwarp();
                        const int sub = lane >> 3, piece = (lane & 7) * 4;
                        const int ty0 = (tile / tiles_w) << (7 - tw_log2), tx0 = (tile % tiles_w) << tw_log2;
#pragma unroll
                        for (int s8 = 0; s8 < 8; ++s8) {
                            const int pl = 4 * s8 + sub;                    // pixel of this warp's 32
                            const int pm = q * 32 + pl;
                            const int py = ty0 + (pm >> tw_log2), px = tx0 + (pm & (tile_w - 1));
                            if (py < H && px < W) {
                                const float* src = stg + pl * 33 + piece;
                                *reinterpret_cast<float4*>(Y + (size_t(py) * W + px) * size_t(Cout) + n0 + c0 + piece) =
                                    make_float4(src[0], src[1], src[2], src[3]);
                            }
                        }
                    } else if (inb) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(Y + rowoff + c0 + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                    }
                }
            }
        }
        if (y_absmax != nullptr) {                                     // max|Y| for the consumer's FP16 scale
            const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(amax));
            if (lane == 0 && wm != 0u) atomicMax(y_absmax, wm);
        }
        tc::tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// style gradient: per 8x16-pixel tile the set of classes whose weight m_k^2 is non-zero somewhere in the tile (bit mask),
// and the largest weight of the whole map (float bits, atomicMax; zeroed by the launcher)
__global__ void __launch_bounds__(TC_BM)
style_tiles_kernel(const float* __restrict__ cls_masks, int num_cls, int H, int W, int tiles_w, uint32_t* __restrict__ tile_active,
                   uint32_t* __restrict__ w_absmax) {
    const int tile = blockIdx.x, t = threadIdx.x;
    const int gy = (tile / tiles_w) * TC_TH + t / TC_TW, gx = (tile % tiles_w) * TC_TW + t % TC_TW;
    const bool inb = gy < H && gx < W;
    uint32_t active = 0u;
    float wmax = 0.f;
    for (int k = 0; k < num_cls; ++k) {
        float w = 0.f;
        if (inb) {
            w = cls_masks ? __ldg(cls_masks + size_t(k) * H * W + size_t(gy) * W + gx) : 1.0f;
            w *= w;
        }
        wmax = fmaxf(wmax, w);
        if (__syncthreads_or(w != 0.f)) active |= 1u << k;
    }
    if (t == 0) tile_active[tile] = active;
    const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(wmax));
    if ((t & 31) == 0 && wm != 0u) atomicMax(w_absmax, wm);
}

// ---------------------------------------------------------------------------------------------------------------
// max|x| of a float32 tensor as float bits (see tc_common.cuh); the slot must have been zeroed.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, size_t n, uint32_t* __restrict__ slot) {
    float m = 0.f;
    const size_t n4 = n / 4;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    if (blockIdx.x == 0 && threadIdx.x < int(n - n4 * 4)) m = fmaxf(m, fabsf(x[n4 * 4 + threadIdx.x]));
    const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if ((threadIdx.x & 31) == 0 && wm != 0u) atomicMax(slot, wm);
}

int launch_absmax(const float* x, size_t n, uint32_t* slot, cudaStream_t st, bool reset) {
    if (reset) ADPST_CUDA_CHECK(cudaMemsetAsync(slot, 0, sizeof(uint32_t), st));
    const size_t want = (n / 4 + 255) / 256, cap = size_t(num_sms()) * 8;
    absmax_kernel<<<unsigned(want < cap ? (want ? want : 1) : cap), 256, 0, st>>>(x, n, slot);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// weight preparation: K-major [tap][N][K] FP16 hi / lo planes, scaled by the power of two derived from max|W|
//   forward : N = Cout, K = Cin :  B[tap][co][ci] = W[tap][ci][co]
//   gradient: N = Cin,  K = Cout:  B[tap][ci][co] = W[8-tap][ci][co]
// ---------------------------------------------------------------------------------------------------------------
__global__ void split_weights_kernel(const float* __restrict__ Wf, __half* __restrict__ hi, __half* __restrict__ lo, int Cin,
                                     int Cout, int gradient, const uint32_t* __restrict__ w_absmax) {
    const size_t total = size_t(9) * Cin * Cout;
    const float sb = tc::pow2f_int(tc::f16_scale_exponent(*w_absmax));
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
        const float t = w * sb;
        const __half h = __float2half_rn(t);
        hi[i] = h;
        lo[i] = __float2half_rn((t - __half2float(h)) * 2048.0f);
    }
}

int prepare_tc_weights(adpst_vgg* h, int i, cudaStream_t st) {
    const int cin = conv_cin(i), cout = conv_cout(i);
    const size_t n = size_t(9) * cin * cout;
    uint32_t* wslot = h->amax + AMAX_WEIGHT + i;
    int rc = launch_absmax(h->wf[i], n, wslot, st);
    if (rc != ADPST_OK) return rc;
    for (int g = 0; g < 2; ++g) {
        ADPST_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->tc_hi[g][i]), n * 2));
        ADPST_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->tc_lo[g][i]), n * 2));
        split_weights_kernel<<<unsigned((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048), 256, 0, st>>>(
            h->wf[i], static_cast<__half*>(h->tc_hi[g][i]), static_cast<__half*>(h->tc_lo[g][i]), cin, cout, g, wslot);
        ADPST_LAUNCH_CHECK();
        // tensor map of the [9*N][K] matrix (K innermost)
        const int N = g ? cin : cout, K = g ? cout : cin;
        const int BN = N >= 128 ? 128 : N;
        const uint64_t dims[2] = {uint64_t(K), uint64_t(9) * N};
        const uint64_t strides[1] = {uint64_t(K) * 2};
        const uint32_t box[2] = {uint32_t(TC_BK), uint32_t(BN)};
        rc = tc::make_tensor_map_f16(&h->tm_hi[g][i], h->tc_hi[g][i], 2, dims, strides, box);
        if (rc != ADPST_OK) return rc;
        rc = tc::make_tensor_map_f16(&h->tm_lo[g][i], h->tc_lo[g][i], 2, dims, strides, box);
        if (rc != ADPST_OK) return rc;
    }
    return ADPST_OK;
}

bool conv_tc_eligible(int Cin, int Cout) { return Cin % TC_BK == 0 && (Cout == 64 || Cout % 128 == 0); }

template <int BN, int MODE>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmBhi, const CUtensorMap& tmBlo, const float* bias, float* Y,
                     const float* seed, const float* mask, int H, int W, int Cin, int Cout, const uint32_t* a_absmax,
                     const uint32_t* b_absmax, uint32_t* y_absmax, cudaStream_t st, const float* cls_masks = nullptr,
                     const uint32_t* tile_active = nullptr, const uint32_t* w_absmax = nullptr, float* pool_out = nullptr,
                     int tw_log2 = 4, int pool_pitch = 0, int pool_xoff = 0) {
    using Cfg = TcCfg<BN, MODE>;
    auto kern = conv3x3_tc_kernel<BN, MODE>;
    ADPST_ONCE_PER_DEVICE(ADPST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)));
    const int tile_w = 1 << tw_log2, tile_h = TC_BM >> tw_log2;
    const int tw = (W + tile_w - 1) / tile_w, th = (H + tile_h - 1) / tile_h;
    const int total = tw * th * (Cout / BN);
    const int grid = total < num_sms() ? total : num_sms();
    kern<<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmBhi, tmBlo, bias, Y, seed, mask, H, W, Cin, Cout, tw, tw * th,
                                                    cls_masks, tile_active, a_absmax, b_absmax, w_absmax, y_absmax, pool_out,
                                                    tw_log2, pool_pitch, pool_xoff);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

// 8 x 16 pixel tiles, or 16 x 8 where that covers the map with fewer of them (narrow maps)
static int pick_tile_shape(int H, int W) {
    const long wide = long((W + 15) / 16) * ((H + 7) / 8), tall = long((W + 7) / 8) * ((H + 15) / 16);
    return tall < wide ? 3 : 4;
}

static int make_act_map(CUtensorMap* tm, const float* X, int H, int W, int C, int tw_log2 = 4) {
    const uint64_t dims[4] = {uint64_t(C), uint64_t(W), uint64_t(H), 1};
    const uint64_t strides[3] = {uint64_t(C) * 4, uint64_t(W) * C * 4, uint64_t(H) * W * C * 4};
    const uint32_t box[4] = {32u, uint32_t((1 << tw_log2) + 2), uint32_t((TC_BM >> tw_log2) + 2), 1};   // pixel tile + 1-pixel halo
    return tc::make_tensor_map_f32(tm, X, 4, dims, strides, box);
}

// X: (H,W,Cin) activation; gradient = 0: conv i forward (bias + ReLU), 1: data gradient of conv i (Cin/Cout are the GEMM's
// K and N, i.e. already swapped for the gradient).  x_absmax: device slot holding max|X| (float bits); y_absmax (may be
// NULL): slot that receives max|Y| (atomicMax; the caller zeroes it).
// pool_out (forward only, may be NULL): receives the 2x2/2 max-pool of Y, (H/2, W/2, Cout), stored with pool_pitch columns per
// row starting at column pool_xoff (W/2 and 0 for a plain tensor).
int launch_conv_tc(adpst_vgg* h, int i, int gradient, const float* X, float* Y, const float* seed, const float* mask, int H,
                   int W, int Cin, int Cout, const uint32_t* x_absmax, uint32_t* y_absmax, float* pool_out, cudaStream_t st,
                   int pool_pitch, int pool_xoff) {
    CUtensorMap tmA;
    const int shape = pick_tile_shape(H, W);
    int rc = make_act_map(&tmA, X, H, W, Cin, shape);
    if (rc != ADPST_OK) return rc;
    const CUtensorMap& bh = h->tm_hi[gradient][i];
    const CUtensorMap& bl = h->tm_lo[gradient][i];
    const float* bias = gradient ? nullptr : h->bias[i];
    const uint32_t* wslot = h->amax + AMAX_WEIGHT + i;
    const int BN = Cout >= 128 ? 128 : Cout;
    if (!gradient) {
        if (BN == 128)
            return launch_tc<128, MODE_FWD>(tmA, bh, bl, bias, Y, seed, mask, H, W, Cin, Cout, x_absmax, wslot, y_absmax, st,
                                            nullptr, nullptr, nullptr, pool_out, shape, pool_pitch, pool_xoff);
        return launch_tc<64, MODE_FWD>(tmA, bh, bl, bias, Y, seed, mask, H, W, Cin, Cout, x_absmax, wslot, y_absmax, st, nullptr,
                                       nullptr, nullptr, pool_out, shape, pool_pitch, pool_xoff);
    }
    if (BN == 128)
        return launch_tc<128, MODE_BWD>(tmA, bh, bl, bias, Y, seed, mask, H, W, Cin, Cout, x_absmax, wslot, y_absmax, st, nullptr,
                                        nullptr, nullptr, nullptr, shape);
    return launch_tc<64, MODE_BWD>(tmA, bh, bl, bias, Y, seed, mask, H, W, Cin, Cout, x_absmax, wslot, y_absmax, st, nullptr,
                                   nullptr, nullptr, nullptr, shape);
}

// ---------------------------------------------------------------------------------------------------------------
// style gradient (components/loss.py:104-137, backward):  dF[px,:] (=|+=) sum_k m_k[px]^2 F[px,:] D_k
// The same kernel with the classes in the role of the filter taps: A = F tile scaled per pixel by m_k^2 in the transform
// warps, B = D_k (symmetric, so K-major == MN-major) as FP16 hi/lo planes scaled by the power of two of *d_absmax.
// Classes whose mask vanishes on the whole 8x16 pixel tile are skipped (exact).
// ---------------------------------------------------------------------------------------------------------------
bool style_tc_eligible(int C) { return C == 64 || C % 128 == 0; }

size_t style_tc_scratch_bytes(int HW) { return size_t(HW) * 4 + 64; }

// per-tile class sets of a (constant) mask stack: tiles[0] = float bits of the largest class weight, tiles[16 + t] = classes
// present in 8x16-pixel tile t.  `tiles` holds style_tc_scratch_bytes(H*W) bytes.
int launch_style_tiles(const float* masks, int K, int H, int W, void* tiles, cudaStream_t st) {
    ADPST_REQUIRE(K >= 1 && K <= TC_MAX_CLASSES, "style gradient: K=%d classes not supported (max %d)", K, TC_MAX_CLASSES);
    const int tw = (W + TC_TW - 1) / TC_TW, th = (H + TC_TH - 1) / TC_TH;
    uint32_t* wslot = static_cast<uint32_t*>(tiles);
    ADPST_CUDA_CHECK(cudaMemsetAsync(wslot, 0, sizeof(uint32_t), st));
    style_tiles_kernel<<<tw * th, TC_BM, 0, st>>>(masks, K, H, W, tw, wslot + 16, wslot);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

// tiles: the buffer launch_style_tiles filled for these masks
int launch_style_dF_tc(const float* F, int H, int W, int C, const float* masks, int K, const void* D_hi, const void* D_lo,
                       const uint32_t* f_absmax, const uint32_t* d_absmax, float* dF, int accumulate, const void* tiles,
                       cudaStream_t st) {
    ADPST_REQUIRE(K >= 1 && K <= TC_MAX_CLASSES, "style gradient: K=%d classes not supported (max %d)", K, TC_MAX_CLASSES);
    CUtensorMap tmA, tmH, tmL;
    int rc = make_act_map(&tmA, F, H, W, C);
    if (rc != ADPST_OK) return rc;
    const int BN = C >= 128 ? 128 : C;
    const uint64_t ddims[2] = {uint64_t(C), uint64_t(K) * C};
    const uint64_t dstr[1] = {uint64_t(C) * 2};
    const uint32_t dbox[2] = {uint32_t(TC_BK), uint32_t(BN)};
    rc = tc::make_tensor_map_f16(&tmH, D_hi, 2, ddims, dstr, dbox);
    if (rc != ADPST_OK) return rc;
    rc = tc::make_tensor_map_f16(&tmL, D_lo, 2, ddims, dstr, dbox);
    if (rc != ADPST_OK) return rc;
    const uint32_t* wslot = static_cast<const uint32_t*>(tiles);
    const uint32_t* tile_active = wslot + 16;
    const float* seed = accumulate ? dF : nullptr;
    if (BN == 128)
        return launch_tc<128, MODE_STYLE>(tmA, tmH, tmL, nullptr, dF, seed, nullptr, H, W, C, C, f_absmax, d_absmax, nullptr, st,
                                          masks, tile_active, wslot);
    return launch_tc<64, MODE_STYLE>(tmA, tmH, tmL, nullptr, dF, seed, nullptr, H, W, C, C, f_absmax, d_absmax, nullptr, st, masks,
                                     tile_active, wslot);
}

}  // namespace adpst
