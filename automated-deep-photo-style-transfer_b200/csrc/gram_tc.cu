// gram_tc.cu -- segmentation-masked Gram matrices on the 5th-generation tensor cores.
//
// Replaces components/loss.py:96-102 (calculate_gram_matrix) for all K classes of one layer:
//     G_k = X_k^T X_k,   X_k = F * m_k  (rows = pixels scaled by the class mask, exactly the reference's formulation).
//
// GEMM view: D[c1, c2] = sum_px X[px, c1] X[px, c2]: the contraction runs over PIXELS, i.e. the operands are needed
// channel-major x pixel ("K-major" with K = pixel) while HBM holds pixel-major x channel.  tcgen05 kind::tf32 does not
// accept MN-major operands (measured: any MN-major tf32 descriptor yields zeros, tests/cuda/umma_mnmajor_probe.cu), so
// the transpose happens on chip: one pipeline stage = a 2 x 16 pixel patch; TMA boxes (32 channels x 16 x 2 pixels) land
// pixel-major, and the transform warps -- which touch every element anyway for the mask scaling and the TF32 hi/lo
// split -- write the K-major, 128B-swizzled X^T tiles the MMA reads (conflict-free: a warp reads one 128-byte row and
// writes 16-byte chunks to 8 different rows).  Diagonal tiles reuse the M-side operand for the N side.
// Only patches where the class mask is non-zero are visited (list built once per layer from the constant masks), so the
// work is ~(1 + boundary fraction) * 2 HW C^2 instead of K * 2 HW C^2.
//
// Precision: 3xTF32 with unbiased hi/lo splits and chunk promotion to registers, as in conv_tc.cu.
// Output: per-(class, split) partial tiles in the workspace; gram_reduce_kernel sums them in float64 (deterministic) and
// mirrors the upper triangle.
#include "tc_common.cuh"
#include "vgg.cuh"

namespace adpst {

constexpr int GM_PH = 2, GM_PW = 16, GM_PX = GM_PH * GM_PW;     // 32 pixels per stage = 4 UMMA K-steps of 8
constexpr int GM_BLK_BYTES = GM_PX * 128;                        // one (32 channel x 32 pixel) box: 4 KB
constexpr int GM_THREADS = 320;
constexpr int GM_CHUNK_ITERS = 2;
constexpr int GM_STAGES = 2;

template <int BN> struct GramCfg {
    static constexpr int RAW_A = 4 * GM_BLK_BYTES;                // landed by TMA: 4 blocks of [32 px][32 ch]
    static constexpr int RAW_B = (BN / 32) * GM_BLK_BYTES;
    static constexpr int A_BYTES = 128 * 128;                     // operand: [128 ch][32 px] K-major, one 128-byte row per channel
    static constexpr int B_BYTES = BN * 128;
    static constexpr int OFF_RAW_B = RAW_A, OFF_AHI = RAW_A + RAW_B, OFF_ALO = OFF_AHI + A_BYTES, OFF_BHI = OFF_ALO + A_BYTES,
                         OFF_BLO = OFF_BHI + B_BYTES;
    static constexpr int STAGE_BYTES = OFF_BLO + B_BYTES;
    static constexpr int SMEM_BYTES = GM_STAGES * STAGE_BYTES + 1024 + 256 + GM_STAGES * GM_PX * 4;
    static constexpr uint32_t TMEM_COLS = 4 * BN;
};

template <int BN>
__global__ void __launch_bounds__(GM_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmF, const float* __restrict__ masks, const int* __restrict__ patch_ids,
               const int* __restrict__ patch_off, float* __restrict__ ws, int H, int W, int C, int splits, int tiles,
               int patches_w) {
    using Cfg = GramCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GM_STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* ready = bars + GM_STAGES;
    uint64_t* empty = bars + 2 * GM_STAGES;
    uint64_t* chunk_full = bars + 3 * GM_STAGES;
    uint64_t* chunk_empty = chunk_full + 2;
    uint64_t* small_full = chunk_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(small_full + 1);
    float* sMw = reinterpret_cast<float*>(bars + 32);                  // [GM_STAGES][32] mask value per pixel of the patch

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

    if (threadIdx.x == 0) {
        for (int s = 0; s < GM_STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
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
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % GM_STAGES, round = it / GM_STAGES;
                tc::mbar_wait(&empty[s], (round & 1) ^ 1);
                const int p = patch_ids[my_begin + it];
                const int y0 = (p / patches_w) * GM_PH, x0 = (p % patches_w) * GM_PW;
                uint8_t* st = smem + s * Cfg::STAGE_BYTES;
                tc::mbar_arrive_expect_tx(&full[s], Cfg::RAW_A + (diag ? 0 : Cfg::RAW_B));
#pragma unroll
                for (int b = 0; b < 4; ++b)          // C == 64: the two channel blocks are loaded twice (rows 64..127 unused)
                    tc::tma_load_4d(st + b * GM_BLK_BYTES, &tmF, &full[s], (tm * 128 + b * 32) % C, x0, y0, 0);
                if (!diag) {
#pragma unroll
                    for (int b = 0; b < BN / 32; ++b)
                        tc::tma_load_4d(st + Cfg::OFF_RAW_B + b * GM_BLK_BYTES, &tmF, &full[s], tn * BN + b * 32, x0, y0, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (warp-uniform loop, one elected lane issues; see conv_tc.cu) =================
        constexpr uint32_t idesc = tc::umma_idesc_tf32(128, BN);
        const uint32_t stage0 = tc::smem_u32(smem);
        const uint64_t d_ahi = tc::umma_desc_kmajor_sw128(stage0 + Cfg::OFF_AHI, 1024);
        const uint64_t d_alo = tc::umma_desc_kmajor_sw128(stage0 + Cfg::OFF_ALO, 1024);
        const uint64_t d_bhi = diag ? d_ahi : tc::umma_desc_kmajor_sw128(stage0 + Cfg::OFF_BHI, 1024);
        const uint64_t d_blo = diag ? d_alo : tc::umma_desc_kmajor_sw128(stage0 + Cfg::OFF_BLO, 1024);
        int s = 0, round = 0;
        for (int it = 0; it < iters; ++it) {
            const int c = it / GM_CHUNK_ITERS, cpos = it - c * GM_CHUNK_ITERS;
            const uint32_t tmem_big = tmem_base + uint32_t(c & 1) * BN;
            if (cpos == 0) tc::mbar_wait(&chunk_empty[c & 1], ((c >> 1) & 1) ^ 1);
            tc::mbar_wait(&ready[s], round & 1);      // the transform warps arrive only after the TMA tile landed
            tc::tcgen05_fence_after();
            const uint64_t soff = uint64_t(uint32_t(s) * uint32_t(Cfg::STAGE_BYTES >> 4));
            if (tc::elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < GM_PX / 8; ++ks) {                  // 8 pixels per MMA = 32 bytes along the K-major row
                    const uint64_t koff = soff + uint64_t(ks * 2);
                    tc::umma_tf32(tmem_small, d_alo + koff, d_bhi + koff, idesc, (it | ks) != 0);
                    tc::umma_tf32(tmem_small, d_ahi + koff, d_blo + koff, idesc, 1);
                    tc::umma_tf32(tmem_big, d_ahi + koff, d_bhi + koff, idesc, (cpos | ks) != 0);
                }
                tc::umma_commit(&empty[s]);
                if (cpos == GM_CHUNK_ITERS - 1 || it == iters - 1) tc::umma_commit(&chunk_full[c & 1]);
            }
            __syncwarp();
            if (++s == GM_STAGES) { s = 0; ++round; }
        }
        if (iters > 0 && tc::elect_one_sync()) tc::umma_commit(small_full);
        __syncwarp();
    } else if (warp < 6) {
        // ================= operand transform: X = m_k * F, split into TF32 hi / lo =================
        const int t = threadIdx.x - 64;                                 // 0..127
        for (int it = 0; it < iters; ++it) {
            const int s = it % GM_STAGES, round = it / GM_STAGES;
            if (t < GM_PX) {
                const int p = patch_ids[my_begin + it];
                const int gy = (p / patches_w) * GM_PH + t / GM_PW, gx = (p % patches_w) * GM_PW + t % GM_PW;
                float m = 0.f;
                if (gy < H && gx < W) m = mk ? __ldg(mk + size_t(gy) * W + gx) : 1.0f;
                sMw[s * GM_PX + t] = m;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");             // the four transform warps only
            tc::mbar_wait(&full[s], round & 1);
            const float* mw = sMw + s * GM_PX;
            uint8_t* st = smem + s * Cfg::STAGE_BYTES;
            const int nsets = diag ? 1 : 2;
            for (int set = 0; set < nsets; ++set) {
                if (set == 1 && t >= BN) break;
                // thread t owns channel row t of the operand: gathers its 32 pixels from the pixel-major landed tile
                const uint8_t* raw = st + (set ? Cfg::OFF_RAW_B : 0) + (t >> 5) * GM_BLK_BYTES;
                uint8_t* ohi = st + (set ? Cfg::OFF_BHI : Cfg::OFF_AHI) + (t >> 3) * 1024 + (t & 7) * 128;
                uint8_t* olo = st + (set ? Cfg::OFF_BLO : Cfg::OFF_ALO) + (t >> 3) * 1024 + (t & 7) * 128;
                const int col = t & 31;
#pragma unroll
                for (int p4 = 0; p4 < GM_PX / 4; ++p4) {
                    float x[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int px = p4 * 4 + j;
                        // landed layout: row = pixel (128 B), 16-byte chunk index XOR (row & 7)
                        const float v = *reinterpret_cast<const float*>(raw + px * 128 + ((((col >> 2) ^ (px & 7)) << 4) | ((col & 3) << 2)));
                        x[j] = v * mw[px];
                    }
                    float4 h, l;
                    h.x = tc::round_tf32(x[0]); l.x = tc::round_tf32(x[0] - h.x);
                    h.y = tc::round_tf32(x[1]); l.y = tc::round_tf32(x[1] - h.y);
                    h.z = tc::round_tf32(x[2]); l.z = tc::round_tf32(x[2] - h.z);
                    h.w = tc::round_tf32(x[3]); l.w = tc::round_tf32(x[3] - h.w);
                    const int chunk = (p4 ^ (t & 7)) << 4;                 // operand layout: row = channel, same XOR swizzle
                    *reinterpret_cast<float4*>(ohi + chunk) = h;
                    *reinterpret_cast<float4*>(olo + chunk) = l;
                }
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&ready[s]);
        }
    } else {
        // ================= drain (chunk promotion) + store of the partial tile =================
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
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(out + c0 + j) =
                        make_float4(acc[c0 + j] + __uint_as_float(v[j]), acc[c0 + j + 1] + __uint_as_float(v[j + 1]),
                                    acc[c0 + j + 2] + __uint_as_float(v[j + 2]), acc[c0 + j + 3] + __uint_as_float(v[j + 3]));
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
                            int H, int W, int C, int K, int splits, cudaStream_t st) {
    using Cfg = GramCfg<BN>;
    auto kern = gram_tc_kernel<BN>;
    static bool configured = false;
    if (!configured) {
        ADPST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int tiles = gram_tc_tiles(C);
    dim3 grid(tiles * (tiles + 1) / 2, K * splits);
    kern<<<grid, GM_THREADS, Cfg::SMEM_BYTES, st>>>(tmF, masks, patch_ids, patch_off, ws, H, W, C, splits, tiles,
                                                    (W + GM_PW - 1) / GM_PW);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

// partial Grams of all classes into ws[(k*splits + s)][C][C] (upper-triangular 128-tiles only)
int launch_gram_tc(const float* F, int H, int W, int C, const float* masks, int K, const int* patch_ids, const int* patch_off,
                   float* ws, int splits, cudaStream_t st) {
    CUtensorMap tmF;
    const uint64_t dims[4] = {uint64_t(C), uint64_t(W), uint64_t(H), 1};
    const uint64_t strides[3] = {uint64_t(C) * 4, uint64_t(W) * C * 4, uint64_t(H) * W * C * 4};
    const uint32_t box[4] = {32u, uint32_t(GM_PW), uint32_t(GM_PH), 1};
    int rc = tc::make_tensor_map_f32(&tmF, F, 4, dims, strides, box);
    if (rc != ADPST_OK) return rc;
    if (C == 64) return launch_gram_tc_t<64>(tmF, masks, patch_ids, patch_off, ws, H, W, C, K, splits, st);
    return launch_gram_tc_t<128>(tmF, masks, patch_ids, patch_off, ws, H, W, C, K, splits, st);
}

}  // namespace adpst
