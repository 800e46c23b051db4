// gram_simt.cu -- segmentation-masked Gram matrices and the style-loss gradient, float32 CUDA-core path.
//
// Replaces   components/loss.py:96-102   calculate_gram_matrix:   G_k = (F * m_k)^T (F * m_k)
//            components/loss.py:104-137  calculate_layer_style_loss and its tape gradient
//
// Both directions are GEMMs whose contraction or row dimension is the pixel index, with the mask folded in as
// a per-pixel weight m_k^2.  Masks are soft only at class boundaries (bilinear resize, loss.py:112-113), so
// (pixel-chunk, class) pairs whose mask is identically zero are skipped; that is exact, not an approximation.
//   forward : G_k[c1,c2] = sum_px m_k[px]^2 F[px,c1] F[px,c2]   (upper-triangular 64x64 tiles, split over pixels,
//             partials reduced in float64 in a fixed order -> deterministic)
//   backward: D_k = 2 s / (C^4 HW^2) (G_k - A_k);   dF[px,:] = sum_k m_k[px]^2 F[px,:] D_k
#include <cuda_fp16.h>
#include "tc_common.cuh"
#include "vgg.cuh"

namespace adpst {

constexpr int GT = 64;        // output tile edge
constexpr int GK = 16;        // contraction chunk
constexpr int GTHREADS = 256;

static inline int gram_tiles(int C) { return (C + GT - 1) / GT; }
static inline int gram_pairs(int C) { const int t = gram_tiles(C); return t * (t + 1) / 2; }

// pixel splits of the tensor-core kernel: enough CTAs for ~2 per SM
static inline int gram_splits_tc(int C, int K) {
    const int t = gram_tc_tiles(C), ctas = t * (t + 1) / 2 * K;
    const int s = (2 * num_sms() + ctas - 1) / ctas;
    return s < 1 ? 1 : s;
}

static inline int gram_splits(int HW, int C, int K) {
    const int ctas = gram_pairs(C) * K;
    int s = (4 * num_sms() + ctas - 1) / ctas;
    const int max_s = (HW + GK * 8 - 1) / (GK * 8);
    if (s > max_s) s = max_s;
    return s < 1 ? 1 : s;
}

// ---------------------------------------------------------------------------------------------
// forward partials: ws[(k * splits + s)][C][C] (only tiles with tn >= tm are written)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GTHREADS)
gram_partial_kernel(const float* __restrict__ F, const float* __restrict__ masks, float* __restrict__ ws, int HW, int C,
                    int K, int splits, int tiles) {
    __shared__ __align__(16) float sA[GK][GT];
    __shared__ __align__(16) float sB[GK][GT];
    __shared__ float sM[GK];
    // decode the (tm <= tn) tile pair
    int pair = blockIdx.x, tm = 0;
    while (pair >= tiles - tm) { pair -= tiles - tm; ++tm; }
    const int tn = tm + pair;
    const int k = blockIdx.y / splits, s = blockIdx.y - k * splits;
    const int chunks = (HW + GK - 1) / GK;
    const int c_begin = int((long long)chunks * s / splits), c_end = int((long long)chunks * (s + 1) / splits);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lrow = tid >> 4, lcol = (tid & 15) * 4;          // loader mapping: 16 rows x 16 float4
    const float* mk = masks ? masks + size_t(k) * HW : nullptr;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int ch = c_begin; ch < c_end; ++ch) {
        const int px0 = ch * GK;
        float m = 0.f;
        if (tid < GK) {
            const int px = px0 + tid;
            m = px < HW ? (mk ? mk[px] : 1.0f) : 0.0f;
            sM[tid] = m * m;
        }
        const int any = __syncthreads_or(m != 0.0f);
        if (!any) continue;                                    // block-uniform: mask is zero on the whole chunk
        {
            const int px = px0 + lrow;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (px < HW) {
                const float* row = F + size_t(px) * C;
                if (tm * GT + lcol < C) a = __ldg(reinterpret_cast<const float4*>(row + tm * GT + lcol));
                if (tn * GT + lcol < C) b = __ldg(reinterpret_cast<const float4*>(row + tn * GT + lcol));
            }
            const float w = sM[lrow];
            b.x *= w; b.y *= w; b.z *= w; b.w *= w;
            *reinterpret_cast<float4*>(&sA[lrow][lcol]) = a;
            *reinterpret_cast<float4*>(&sB[lrow][lcol]) = b;
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < GK; ++p) {
            const float4 a = *reinterpret_cast<const float4*>(&sA[p][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&sB[p][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = ws + size_t(blockIdx.y) * C * C;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = tm * GT + ty * 4 + i, c = tn * GT + tx * 4;
        if (r < C && c < C) *reinterpret_cast<float4*>(out + size_t(r) * C + c) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// G[k][r][c] = sum_s ws[k][s][r][c] for the 64x64 tiles on or above the diagonal (the only ones the Gram kernels write),
// float64 accumulation in a fixed order; tiles above the diagonal are mirrored into the lower triangle through a
// shared-memory transpose, so that every global access is a coalesced row segment.
// grid = (C/32, C/32, K), block = (32, 8).
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const float* __restrict__ ws, float* __restrict__ G, int C, int K, int splits) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32, k = blockIdx.z;
    const int tr = r0 / GT, tcx = c0 / GT;
    if (tr > tcx) return;                                       // filled by the mirror of (tcx, tr)
    const float* src = ws + (size_t(k) * splits) * C * C;
    float* dst = G + size_t(k) * C * C;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + threadIdx.y + 8 * i, c = c0 + threadIdx.x;
        double a = 0.0;
        const float* p = src + size_t(r) * C + c;
        for (int s = 0; s < splits; ++s) a += double(p[size_t(s) * C * C]);
        const float v = float(a);
        dst[size_t(r) * C + c] = v;
        tile[threadIdx.y + 8 * i][threadIdx.x] = v;
    }
    if (tr == tcx) return;                                      // a diagonal tile holds both of its halves already
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + threadIdx.y + 8 * i, r = r0 + threadIdx.x;
        dst[size_t(c) * C + r] = tile[threadIdx.x][threadIdx.y + 8 * i];
    }
}

// ---------------------------------------------------------------------------------------------
// backward step 1: D_k = coef (G_k - A_k),  loss += loss_coef * sum (G_k - A_k)^2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
style_diff_kernel(const float* __restrict__ G, const float* __restrict__ A, float* __restrict__ D, size_t n, double coef,
                  double loss_coef, double* __restrict__ loss, uint32_t* __restrict__ d_absmax) {
    __shared__ double red[32];
    double acc = 0.0;
    float amax = 0.f;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const double d = double(G[i]) - double(A[i]);
        acc += d * d;
        const float v = float(coef * d);
        D[i] = v;
        amax = fmaxf(amax, fabsf(v));
    }
    const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(amax));     // for the FP16 scale (tc_common.cuh)
    if ((threadIdx.x & 31) == 0 && wm != 0u) atomicMax(d_absmax, wm);
    acc = block_sum<double>(acc, red);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, acc * loss_coef);
}

// FP16 hi / lo planes of D for the tensor-core kernel, scaled by the power of two of max|D|
__global__ void __launch_bounds__(256)
style_split_kernel(const float* __restrict__ D, __half* __restrict__ hi, __half* __restrict__ lo, size_t n,
                   const uint32_t* __restrict__ d_absmax) {
    const float sd = tc::pow2f_int(tc::f16_scale_exponent(*d_absmax));
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const float t = D[i] * sd;
        const __half h = __float2half_rn(t);
        hi[i] = h;
        lo[i] = __float2half_rn((t - __half2float(h)) * 2048.0f);
    }
}

// ---------------------------------------------------------------------------------------------
// backward step 2: dF[px, c] (=|+=) sum_k m_k[px]^2 sum_c' F[px,c'] D_k[c',c]
//   CTA tile 64 px x 64 c, 4x4 per thread, contraction over (k, c') in chunks of 16.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GTHREADS)
style_dF_kernel(const float* __restrict__ F, const float* __restrict__ masks, const float* __restrict__ D,
                float* __restrict__ dF, int HW, int C, int K, int accumulate) {
    __shared__ __align__(16) float sA[GK][GT + 4];     // [c'][px], scaled by m_k^2
    __shared__ __align__(16) float sB[GK][GT];         // [c'][c]
    __shared__ float sM[GT];
    const int px0 = blockIdx.x * GT, c0 = blockIdx.y * GT;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k = 0; k < K; ++k) {
        float m = 0.f;
        if (tid < GT) {
            const int px = px0 + tid;
            m = px < HW ? (masks ? masks[size_t(k) * HW + px] : 1.0f) : 0.0f;
            sM[tid] = m * m;
        }
        const int any = __syncthreads_or(m != 0.0f);
        if (!any) continue;
        const float* Dk = D + size_t(k) * C * C;
        for (int cc = 0; cc < C; cc += GK) {
            {   // A: 64 px x 16 c' -> transposed
                const int lpx = tid >> 2, lc4 = (tid & 3) * 4;
                const int px = px0 + lpx;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                if (px < HW) a = __ldg(reinterpret_cast<const float4*>(F + size_t(px) * C + cc + lc4));
                const float w = sM[lpx];
                sA[lc4 + 0][lpx] = a.x * w; sA[lc4 + 1][lpx] = a.y * w; sA[lc4 + 2][lpx] = a.z * w; sA[lc4 + 3][lpx] = a.w * w;
                // B: 16 c' x 64 c
                const int lr = tid >> 4, lc = (tid & 15) * 4;
                *reinterpret_cast<float4*>(&sB[lr][lc]) =
                    __ldg(reinterpret_cast<const float4*>(Dk + size_t(cc + lr) * C + c0 + lc));
            }
            __syncthreads();
#pragma unroll
            for (int p = 0; p < GK; ++p) {
                const float4 a = *reinterpret_cast<const float4*>(&sA[p][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&sB[p][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int px = px0 + ty * 4 + i;
        if (px >= HW) continue;
        float4* o = reinterpret_cast<float4*>(dF + size_t(px) * C + c0 + tx * 4);
        float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (accumulate) { const float4 old = *o; v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
        *o = v;
    }
}

}  // namespace adpst

extern "C" {

size_t adpst_gram_workspace_bytes(int HW, int C, int K) {
    using namespace adpst;
    if (HW <= 0 || C <= 0 || K <= 0) return 0;
    int splits = gram_splits(HW, C, K);
    if (gram_tc_eligible(C) && gram_splits_tc(C, K) > splits) splits = gram_splits_tc(C, K);
    const size_t partials = size_t(K) * splits * C * C * sizeof(float) + 64;       // + the Gram kernel's two scale slots
    // D (fp32), D_hi, D_lo (fp16), two scale slots, then the tensor-core style gradient's scratch
    const size_t dmat = size_t(2) * K * C * C * sizeof(float) + 64 + style_tc_scratch_bytes(HW);
    return partials > dmat ? partials : dmat;
}

int adpst_gram_masked(const float* F_dev, int h, int w, int C, const float* masks_dev, int K, const int* patch_ids_dev,
                      const int* patch_off_dev, float* G_dev, int path, const uint32_t* F_absmax_dev,
                      const uint32_t* masks_absmax_dev, void* workspace_dev, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(F_dev && G_dev && workspace_dev, "gram_masked: NULL argument");
    ADPST_REQUIRE(h > 0 && w > 0 && K > 0, "gram_masked: empty input");
    ADPST_REQUIRE(C > 0 && C % GT == 0, "gram_masked: C=%d must be a multiple of %d", C, GT);
    ADPST_REQUIRE(masks_dev || K == 1, "gram_masked: K=%d needs masks", K);
    cudaStream_t st = as_stream(stream);
    const int HW = h * w;
    int splits;
    if (path == CONV_PATH_TENSOR && patch_ids_dev && patch_off_dev && gram_tc_eligible(C)) {
        splits = gram_splits_tc(C, K);
        // two scale slots behind the partial tiles: max|mask| and, if the caller has no slot for it, max|F|
        uint32_t* slots = reinterpret_cast<uint32_t*>(static_cast<float*>(workspace_dev) + size_t(K) * splits * C * C);
        int rc = ADPST_OK;
        if (masks_dev != nullptr && masks_absmax_dev == nullptr) {
            rc = launch_absmax(masks_dev, size_t(K) * HW, slots, st);
            masks_absmax_dev = slots;
        }
        if (rc == ADPST_OK && F_absmax_dev == nullptr) {
            rc = launch_absmax(F_dev, size_t(HW) * C, slots + 1, st);
            F_absmax_dev = slots + 1;
        }
        if (rc == ADPST_OK)
            rc = launch_gram_tc(F_dev, h, w, C, masks_dev, K, patch_ids_dev, patch_off_dev, static_cast<float*>(workspace_dev),
                                splits, F_absmax_dev, masks_dev ? masks_absmax_dev : nullptr, st);
        if (rc != ADPST_OK) return rc;
    } else {
        splits = gram_splits(HW, C, K);
        const int tiles = gram_tiles(C);
        dim3 grid(gram_pairs(C), K * splits);
        gram_partial_kernel<<<grid, GTHREADS, 0, st>>>(F_dev, masks_dev, static_cast<float*>(workspace_dev), HW, C, K, splits,
                                                       tiles);
        ADPST_LAUNCH_CHECK();
    }
    gram_reduce_kernel<<<dim3(C / 32, C / 32, K), dim3(32, 8), 0, st>>>(static_cast<const float*>(workspace_dev), G_dev, C, K,
                                                                        splits);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

int adpst_style_layer_backward(const float* F_dev, int h, int w, int C, const float* masks_dev, int K, const float* G_dev,
                               const float* A_dev, double loss_scale, double grad_scale, double* loss_dev, float* dF_dev,
                               int accumulate, int path, double hw_norm, const uint32_t* F_absmax_dev, const void* tiles_dev,
                               void* workspace_dev, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(F_dev && G_dev && A_dev && workspace_dev, "style_layer_backward: NULL argument");
    ADPST_REQUIRE(h > 0 && w > 0 && K > 0, "style_layer_backward: empty input");
    const int HW = h * w;
    ADPST_REQUIRE(C > 0 && C % GT == 0, "style_layer_backward: C=%d must be a multiple of %d", C, GT);
    ADPST_REQUIRE(masks_dev || K == 1, "style_layer_backward: K=%d needs masks", K);
    cudaStream_t st = as_stream(stream);
    // spatially tiled runs normalise by the pixel count of the WHOLE image (loss.py:123), not of the local tile
    const double hwn = hw_norm > 0.0 ? hw_norm : double(HW);
    const double c2 = double(C) * double(C), hw2 = hwn * hwn;
    const double coef = 2.0 * grad_scale / (c2 * c2 * hw2);        // D_k = coef (G_k - A_k)
    const double loss_coef = loss_scale / (2.0 * c2 * c2 * hw2);   // L = sum_k sum (G_k - A_k)^2 / (2 C^4 HW^2)
    const size_t n = size_t(K) * C * C;
    // workspace: D (n float32) | D_hi (n fp16) | D_lo (n fp16) | slot: max|D| | slot: max|F| (when the caller has none)
    float* D = static_cast<float*>(workspace_dev);
    __half* Dhi = reinterpret_cast<__half*>(D + n);
    __half* Dlo = Dhi + n;
    uint32_t* slots = reinterpret_cast<uint32_t*>(D + 2 * n);
    const unsigned dgrid = unsigned((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
    ADPST_CUDA_CHECK(cudaMemsetAsync(slots, 0, sizeof(uint32_t), st));
    style_diff_kernel<<<dgrid, 256, 0, st>>>(G_dev, A_dev, D, n, coef, loss_coef, loss_dev, slots);
    ADPST_LAUNCH_CHECK();
    if (dF_dev) {
        if (path == CONV_PATH_TENSOR && style_tc_eligible(C) && K <= 32) {
            style_split_kernel<<<dgrid, 256, 0, st>>>(D, Dhi, Dlo, n, slots);
            ADPST_LAUNCH_CHECK();
            if (F_absmax_dev == nullptr) {
                int rc = launch_absmax(F_dev, size_t(HW) * C, slots + 1, st);
                if (rc != ADPST_OK) return rc;
                F_absmax_dev = slots + 1;
            }
            if (tiles_dev == nullptr) {                      // the caller did not precompute the class sets of its masks
                int rc = launch_style_tiles(masks_dev, K, h, w, slots + 16, st);
                if (rc != ADPST_OK) return rc;
                tiles_dev = slots + 16;
            }
            return launch_style_dF_tc(F_dev, h, w, C, masks_dev, K, Dhi, Dlo, F_absmax_dev, slots, dF_dev, accumulate, tiles_dev, st);
        }
        dim3 grid((HW + GT - 1) / GT, C / GT);
        style_dF_kernel<<<grid, GTHREADS, 0, st>>>(F_dev, masks_dev, D, dF_dev, HW, C, K, accumulate);
        ADPST_LAUNCH_CHECK();
    }
    return ADPST_OK;
}

size_t adpst_style_tiles_bytes(int HW) { return HW > 0 ? adpst::style_tc_scratch_bytes(HW) : 0; }

int adpst_style_tiles(const float* masks_dev, int K, int h, int w, void* tiles_dev, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(tiles_dev && h > 0 && w > 0 && K > 0, "style_tiles: bad argument");
    ADPST_REQUIRE(masks_dev || K == 1, "style_tiles: K=%d needs masks", K);
    return launch_style_tiles(masks_dev, K, h, w, tiles_dev, as_stream(stream));
}

}  // extern "C"
