// vgg_simt.cu -- VGG19 (block1_conv1 .. block5_conv1) forward and data-gradient, float32 CUDA-core path.
//
// Replaces   components/VGG19/model.py:27-41  (x255, caffe preprocess, Keras VGG19 convs/pools, post-ReLU taps)
//            style_transfer.py:341            (tape.gradient through the extractor; weights frozen model.py:11,25)
//
// This is the exact-float32 path: every product and sum is IEEE float32, so it holds the 1e-5 parity bar by
// construction.  The tcgen05 (3xFP16) implicit-GEMM kernels in conv_tc.cu are validated against it.
//
// Layout: activations NHWC (N = 1); forward weights [tap][Cin][Cout] (= Keras HWIO), gradient weights
// [tap'][Cout][Cin] with the taps flipped, so the data gradient is the same 3x3 SAME convolution.
#include "vgg.cuh"

namespace adpst {

// pool j sits after conv {1, 3, 7, 11}
static const int kPoolAfter[ADPST_VGG_NUM_POOL] = {1, 3, 7, 11};
static inline int pools_before(int conv) { return conv >= 12 ? 4 : conv >= 8 ? 3 : conv >= 4 ? 2 : conv >= 2 ? 1 : 0; }

// ---------------------------------------------------------------------------------------------
// 3x3 SAME convolution as an implicit GEMM on CUDA cores.
//   CTA tile: 8 x 16 pixels x 64 output channels, 128 threads, 8 px x 8 ch per thread, Cin in chunks of 8.
// MODE_FWD : y = relu(acc + bias)
// MODE_BWD : g = acc (+ seed);  y = mask_src > 0 ? g : 0      (seed / mask_src optional)
// PRE      : the input is the [0,1] RGB image; 255*x - mean with the channel flip is applied on load
//            (model.py:28-29); zero padding applies to the preprocessed tensor.
// ---------------------------------------------------------------------------------------------
constexpr int CT_H = 8, CT_W = 16, CT_N = 64, CT_K = 8, CT_THREADS = 128, CT_XP = 20;

template <int MODE, bool PRE>
__global__ void __launch_bounds__(CT_THREADS)
conv3x3_simt_kernel(const float* __restrict__ X, const float* __restrict__ Wt, const float* __restrict__ bias,
                    float* __restrict__ Y, const float* __restrict__ seed, const float* __restrict__ mask_src,
                    int H, int W, int Cin, int Cout, int tiles_w, uint32_t* __restrict__ y_absmax) {
    constexpr int KC = PRE ? 3 : CT_K;          // input channels per chunk: the RGB image is not padded to 8
    __shared__ __align__(16) float sX[KC][CT_H + 2][CT_XP];
    __shared__ __align__(16) float sW[9][KC][CT_N];

    const int tile = blockIdx.x;
    const int y0 = (tile / tiles_w) * CT_H, x0 = (tile % tiles_w) * CT_W;
    const int n0 = blockIdx.y * CT_N;
    const int tid = threadIdx.x;
    const int tx = tid & 7, ty = tid >> 3;
    const int row = ty >> 1, wseg = (ty & 1) * 8;

    float acc[8][8];
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[p][n] = 0.0f;

    for (int ci0 = 0; ci0 < Cin; ci0 += KC) {
        // ---- stage the input halo tile, transposed to [ci][row][col]
        if (PRE) {
            for (int px = tid; px < (CT_H + 2) * (CT_W + 2); px += CT_THREADS) {
                const int r = px / (CT_W + 2), c = px - r * (CT_W + 2);
                const int gy = y0 - 1 + r, gx = x0 - 1 + c;
                float v0 = 0.f, v1 = 0.f, v2 = 0.f;
                if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                    const float* q = X + (size_t(gy) * W + gx) * 3;
                    v0 = 255.0f * q[2] - 103.939f;      // B
                    v1 = 255.0f * q[1] - 116.779f;      // G
                    v2 = 255.0f * q[0] - 123.68f;       // R
                }
                sX[0][r][c] = v0; sX[1][r][c] = v1; sX[2][r][c] = v2;
            }
        } else {
            for (int idx = tid; idx < (CT_H + 2) * (CT_W + 2) * 2; idx += CT_THREADS) {
                const int px = idx >> 1, half = idx & 1;
                const int r = px / (CT_W + 2), c = px - r * (CT_W + 2);
                const int gy = y0 - 1 + r, gx = x0 - 1 + c;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gy >= 0 && gy < H && gx >= 0 && gx < W)
                    v = __ldg(reinterpret_cast<const float4*>(X + (size_t(gy) * W + gx) * Cin + ci0 + half * 4));
                if (!PRE) {
                    sX[half * 4 + 0][r][c] = v.x; sX[half * 4 + 1][r][c] = v.y;
                    sX[half * 4 + 2][r][c] = v.z; sX[half * 4 + 3][r][c] = v.w;
                }
            }
        }
        // ---- stage the weights of this chunk: [tap][ci][64]
        for (int idx = tid; idx < 9 * KC * (CT_N / 4); idx += CT_THREADS) {
            const int tap = idx / (KC * (CT_N / 4));
            const int rem = idx - tap * (KC * (CT_N / 4));
            const int ci = rem / (CT_N / 4), n4 = rem - ci * (CT_N / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ci0 + ci < Cin)
                v = __ldg(reinterpret_cast<const float4*>(Wt + (size_t(tap) * Cin + ci0 + ci) * Cout + n0 + n4 * 4));
            *reinterpret_cast<float4*>(&sW[tap][ci][n4 * 4]) = v;
        }
        __syncthreads();

#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll(PRE ? 3 : 2)
            for (int ci = 0; ci < KC; ++ci) {
                float xv[10];
                const float* xr = &sX[ci][row + kh][wseg];
                const float4 a = *reinterpret_cast<const float4*>(xr);
                const float4 b = *reinterpret_cast<const float4*>(xr + 4);
                const float2 c = *reinterpret_cast<const float2*>(xr + 8);
                xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
                xv[8] = c.x; xv[9] = c.y;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&sW[kh * 3 + kw][ci][tx * 8]);
                    const float4 w1 = *reinterpret_cast<const float4*>(&sW[kh * 3 + kw][ci][tx * 8 + 4]);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int p = 0; p < 8; ++p)
#pragma unroll
                        for (int n = 0; n < 8; ++n) acc[p][n] = fmaf(xv[p + kw], wv[n], acc[p][n]);
                }
            }
        }
        __syncthreads();
    }

    // ---- epilogue
    const int gy = y0 + row;
    const int nb = n0 + tx * 8;
    float amax = 0.f;
    float bv[8];
    if (MODE == MODE_FWD) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + nb));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + nb + 4));
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const int gx = x0 + wseg + p;
        if (gx >= W || gy >= H) continue;
        const size_t o = (size_t(gy) * W + gx) * Cout + nb;
        float r[8];
        if (MODE == MODE_FWD) {
#pragma unroll
            for (int n = 0; n < 8; ++n) r[n] = fmaxf(acc[p][n] + bv[n], 0.0f);
        } else {
#pragma unroll
            for (int n = 0; n < 8; ++n) r[n] = acc[p][n];
            if (seed != nullptr) {
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(seed + o));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(seed + o + 4));
                r[0] += s0.x; r[1] += s0.y; r[2] += s0.z; r[3] += s0.w; r[4] += s1.x; r[5] += s1.y; r[6] += s1.z; r[7] += s1.w;
            }
            if (mask_src != nullptr) {
                const float4 m0 = __ldg(reinterpret_cast<const float4*>(mask_src + o));
                const float4 m1 = __ldg(reinterpret_cast<const float4*>(mask_src + o + 4));
                const float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                for (int n = 0; n < 8; ++n) r[n] = mv[n] > 0.0f ? r[n] : 0.0f;
            }
        }
#pragma unroll
        for (int n = 0; n < 8; ++n) amax = fmaxf(amax, fabsf(r[n]));
        *reinterpret_cast<float4*>(Y + o) = make_float4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<float4*>(Y + o + 4) = make_float4(r[4], r[5], r[6], r[7]);
    }
    if (y_absmax != nullptr) {                  // max|Y| for the tensor-core consumer's FP16 scale (tc_common.cuh)
        const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(amax));
        if ((tid & 31) == 0 && wm != 0u) atomicMax(y_absmax, wm);
    }
}

// ---------------------------------------------------------------------------------------------
// data gradient of block1_conv1 back to the [0,1] RGB image (Cout' = 3: too thin for the tiled kernel)
//   dImg[p, rgb] = 255 * sum_{tap, co} dPre[p + tap - 1, co] * W[2-kh][2-kw][ci = 2 - rgb][co]
// Wg: [tap'][3 (rgb)][64] already flipped / permuted / scaled by 255.
// ---------------------------------------------------------------------------------------------
// 16 x 32 pixel tile per CTA, 128 threads; thread (warp w, lane l) owns the four pixels (4w .. 4w+3, l).  The 18 x 34 halo of
// dPre is staged in shared memory 16 channels at a time (49 KB: four CTAs per SM), so every dPre value is read from
// HBM/L2 ~1.2 times instead of 9.  Register blocking over four rows matters because the kernel is bound by shared-memory
// wavefronts, not FMAs: per (kw, 4 channels) a thread issues 6 dPre loads + 9 broadcast weight loads for 144 FMAs
// (one pixel per thread would need 7 loads for 12).
constexpr int IG_TH = 16, IG_TW = 32, IG_CH = 16, IG_ROWS = 4;
constexpr int IG_THREADS = (IG_TH / IG_ROWS) * IG_TW;
constexpr int IG_PITCH = IG_CH + 4;                   // floats per staged pixel (16-byte aligned, bank-skewed)
constexpr int IG_SMEM = (IG_TH + 2) * (IG_TW + 2) * IG_PITCH * 4 + 9 * 3 * 64 * 4;

__global__ void __launch_bounds__(IG_THREADS)
conv1_dgrad_image_kernel(const float* __restrict__ dPre, const float* __restrict__ Wg, float* __restrict__ dImg, int H,
                         int W) {
    extern __shared__ __align__(16) float ig_smem[];
    float* sD = ig_smem;                                              // [(TH+2)*(TW+2)][IG_PITCH]
    float* sW = ig_smem + (IG_TH + 2) * (IG_TW + 2) * IG_PITCH;       // [tap][rgb][64]
    for (int i = threadIdx.x; i < 9 * 3 * 64; i += blockDim.x) sW[i] = Wg[i];
    const int x0 = blockIdx.x * IG_TW, y0 = blockIdx.y * IG_TH;
    const int tx = threadIdx.x & 31, ty = (threadIdx.x >> 5) * IG_ROWS;
    float acc[IG_ROWS][3];
#pragma unroll
    for (int r = 0; r < IG_ROWS; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 0.f;
    for (int part = 0; part < 64 / IG_CH; ++part) {
        __syncthreads();
        for (int i = threadIdx.x; i < (IG_TH + 2) * (IG_TW + 2) * (IG_CH / 4); i += blockDim.x) {
            const int px = i / (IG_CH / 4), c4 = i - px * (IG_CH / 4);
            const int r = px / (IG_TW + 2), c = px - r * (IG_TW + 2);
            const int gy = y0 - 1 + r, gx = x0 - 1 + c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < H && gx >= 0 && gx < W)
                v = __ldg(reinterpret_cast<const float4*>(dPre + (size_t(gy) * W + gx) * 64 + part * IG_CH) + c4);
            *reinterpret_cast<float4*>(sD + px * IG_PITCH + c4 * 4) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
            for (int c4 = 0; c4 < IG_CH / 4; ++c4) {
                float4 d[IG_ROWS + 2];
#pragma unroll
                for (int j = 0; j < IG_ROWS + 2; ++j)
                    d[j] = *reinterpret_cast<const float4*>(sD + ((ty + j) * (IG_TW + 2) + tx + kw) * IG_PITCH + c4 * 4);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const float* w = sW + (kh * 3 + kw) * 192 + part * IG_CH + c4 * 4;
                    const float4 w0 = *reinterpret_cast<const float4*>(w);
                    const float4 w1 = *reinterpret_cast<const float4*>(w + 64);
                    const float4 w2 = *reinterpret_cast<const float4*>(w + 128);
#pragma unroll
                    for (int r = 0; r < IG_ROWS; ++r) {
                        const float4 v = d[r + kh];
                        acc[r][0] = fmaf(v.x, w0.x, acc[r][0]); acc[r][0] = fmaf(v.y, w0.y, acc[r][0]);
                        acc[r][0] = fmaf(v.z, w0.z, acc[r][0]); acc[r][0] = fmaf(v.w, w0.w, acc[r][0]);
                        acc[r][1] = fmaf(v.x, w1.x, acc[r][1]); acc[r][1] = fmaf(v.y, w1.y, acc[r][1]);
                        acc[r][1] = fmaf(v.z, w1.z, acc[r][1]); acc[r][1] = fmaf(v.w, w1.w, acc[r][1]);
                        acc[r][2] = fmaf(v.x, w2.x, acc[r][2]); acc[r][2] = fmaf(v.y, w2.y, acc[r][2]);
                        acc[r][2] = fmaf(v.z, w2.z, acc[r][2]); acc[r][2] = fmaf(v.w, w2.w, acc[r][2]);
                    }
                }
            }
        }
    }
    const int gx = x0 + tx;
#pragma unroll
    for (int r = 0; r < IG_ROWS; ++r) {
        const int gy = y0 + ty + r;
        if (gx < W && gy < H) {
            float* o = dImg + (size_t(gy) * W + gx) * 3;
            o[0] = acc[r][0]; o[1] = acc[r][1]; o[2] = acc[r][2];
        }
    }
}

// max|.| of what a kernel wrote, as float bits, for the FP16 scale of the tensor-core consumer (tc_common.cuh): one
// atomicMax per warp.  Every thread of the (converged) warp must call it.
__device__ __forceinline__ void record_absmax(float amax, uint32_t* __restrict__ slot) {
    if (slot == nullptr) return;
    const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(amax));
    if ((threadIdx.x & 31) == 0 && wm != 0u) atomicMax(slot, wm);
}

// ---------------------------------------------------------------------------------------------
// 2x2/2 VALID max-pool, forward and (fused with the ReLU mask and an optional loss seed) backward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool2_kernel(const float* __restrict__ X, float* __restrict__ P, int H, int W, int C, int p_pitch, int p_xoff) {
    const int Hp = H / 2, Wp = W / 2, C4 = C / 4;
    const size_t total = size_t(Hp) * Wp * C4;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int c4 = int(i % C4);
        const size_t pp = i / C4;
        const int px = int(pp % Wp), py = int(pp / Wp);
        const float4* b = reinterpret_cast<const float4*>(X + (size_t(2 * py) * W + 2 * px) * C) + c4;
        const float4 v00 = __ldg(b), v01 = __ldg(b + C4);
        const float4 v10 = __ldg(b + size_t(W) * C4), v11 = __ldg(b + size_t(W) * C4 + C4);
        float4 m;
        m.x = fmaxf(fmaxf(v00.x, v01.x), fmaxf(v10.x, v11.x));
        m.y = fmaxf(fmaxf(v00.y, v01.y), fmaxf(v10.y, v11.y));
        m.z = fmaxf(fmaxf(v00.z, v01.z), fmaxf(v10.z, v11.z));
        m.w = fmaxf(fmaxf(v00.w, v01.w), fmaxf(v10.w, v11.w));
        reinterpret_cast<float4*>(P)[(size_t(py) * p_pitch + px + p_xoff) * C4 + c4] = m;     // (p_pitch = Wp, p_xoff = 0: P[i])
    }
}

// dPre[pos] = Y[pos] > 0 ? ((pos is the first max of its window ? dP[window] : 0) + seed[pos]) : 0
// (Tried in round 2: arg-max nibbles written by the forward pass's fused-pool epilogue, so that this kernel reads dP and the
// codes only.  It took 159 us instead of 244 us over the four pools at 1024^2, but the two extra shuffles per element in the
// convolution epilogue cost more than that (block1_conv2 forward 333 -> 417 us): not kept.)
// One thread per (2x2 cell, 4 channels): every Y / seed / dPre element is touched exactly once.  Cells cut by an odd H or W
// are not pooling windows (VALID pooling); their pixels only see the seed.
__global__ void __launch_bounds__(256)
unpool_relu_kernel(const float* __restrict__ Y, const float* __restrict__ dP, const float* __restrict__ seed,
                   float* __restrict__ dPre, int H, int W, int C, uint32_t* __restrict__ out_absmax, int dp_pitch, int dp_xoff) {
    // dP is stored with dp_pitch columns per row and the window of cell cx at column cx + dp_xoff (a strip of a tiled run keeps
    // a wider halo on the pooled side; dp_pitch = W / 2, dp_xoff = 0 otherwise)
    const int Hp = H / 2, Wp = W / 2, Hc = (H + 1) / 2, Wc = (W + 1) / 2, C4 = C / 4;
    const size_t total = size_t(Hc) * Wc * C4;
    float amax = 0.f;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int c4 = int(i % C4);
        const size_t cell = i / C4;
        const int cx = int(cell % Wc), cy = int(cell / Wc);
        const bool window = cy < Hp && cx < Wp;
        float4 v[4], g[4];
        bool ok[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int y = 2 * cy + (j >> 1), x = 2 * cx + (j & 1);
            ok[j] = y < H && x < W;
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[j]) v[j] = __ldg(reinterpret_cast<const float4*>(Y + (size_t(y) * W + x) * C) + c4);
        }
        if (window) {
            const float4 d = __ldg(reinterpret_cast<const float4*>(dP + (size_t(cy) * dp_pitch + cx + dp_xoff) * C) + c4);
            const float* vf = reinterpret_cast<const float*>(v);
            float* gf = reinterpret_cast<float*>(g);
            const float df[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int arg = 0;
                float best = vf[k];
#pragma unroll
                for (int j = 1; j < 4; ++j)
                    if (vf[j * 4 + k] > best) { best = vf[j * 4 + k]; arg = j; }      // first maximum wins, as in TF
#pragma unroll
                for (int j = 0; j < 4; ++j) gf[j * 4 + k] = (arg == j) ? df[k] : 0.0f;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!ok[j]) continue;
            const int y = 2 * cy + (j >> 1), x = 2 * cx + (j & 1);
            const size_t o = (size_t(y) * W + x) * C4 + c4;
            float4 r = g[j];
            if (seed != nullptr) {
                const float4 sd = __ldg(reinterpret_cast<const float4*>(seed) + o);
                r.x += sd.x; r.y += sd.y; r.z += sd.z; r.w += sd.w;
            }
            r.x = v[j].x > 0.f ? r.x : 0.f; r.y = v[j].y > 0.f ? r.y : 0.f;
            r.z = v[j].z > 0.f ? r.z : 0.f; r.w = v[j].w > 0.f ? r.w : 0.f;
            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(r.x), fabsf(r.y))), fmaxf(fabsf(r.z), fabsf(r.w)));
            reinterpret_cast<float4*>(dPre)[o] = r;
        }
    }
    record_absmax(amax, out_absmax);
}

// dPre = Y > 0 ? seed : 0     (top of the chain)
__global__ void __launch_bounds__(256)
relu_mask_kernel(const float* __restrict__ Y, const float* __restrict__ seed, float* __restrict__ dPre, size_t n4,
                 uint32_t* __restrict__ out_absmax) {
    float amax = 0.f;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
        const float4 y = __ldg(reinterpret_cast<const float4*>(Y) + i);
        float4 s = __ldg(reinterpret_cast<const float4*>(seed) + i);
        s.x = y.x > 0.f ? s.x : 0.f; s.y = y.y > 0.f ? s.y : 0.f; s.z = y.z > 0.f ? s.z : 0.f; s.w = y.w > 0.f ? s.w : 0.f;
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(s.x), fabsf(s.y))), fmaxf(fabsf(s.z), fabsf(s.w)));
        reinterpret_cast<float4*>(dPre)[i] = s;
    }
    record_absmax(amax, out_absmax);
}

// gradient weights: Wb[tap'][co][ci] = W[8 - tap'][ci][co]
__global__ void flip_transpose_weights_kernel(const float* __restrict__ Wf, float* __restrict__ Wb, int Cin, int Cout) {
    const size_t total = size_t(9) * Cin * Cout;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const int ci = int(i % Cin);
        const size_t r = i / Cin;
        const int co = int(r % Cout), tap = int(r / Cout);
        Wb[i] = Wf[(size_t(8 - tap) * Cin + ci) * Cout + co];
    }
}

// image-gradient weights of conv 0: Wg[tap'][rgb][co] = 255 * W[8 - tap'][ci = 2 - rgb][co]
__global__ void conv1_image_weights_kernel(const float* __restrict__ Wf, float* __restrict__ Wg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * 3 * 64) return;
    const int co = i % 64, rgb = (i / 64) % 3, tap = i / 192;
    Wg[i] = 255.0f * Wf[(size_t(8 - tap) * 3 + (2 - rgb)) * 64 + co];
}

static inline unsigned stream_grid(size_t items, int threads = 256) {
    const size_t want = (items + threads - 1) / threads, cap = size_t(num_sms()) * 16;
    return unsigned(want < cap ? (want ? want : 1) : cap);
}

}  // namespace adpst

// ---------------------------------------------------------------------------------------------
// handle + orchestration
// ---------------------------------------------------------------------------------------------
namespace adpst {

static void layer_hw(int conv, int H, int W, int* h, int* w) {
    const int p = pools_before(conv);
    *h = H >> p;     // floor division by 2, p times (VALID pooling)
    *w = W >> p;
}

// Column geometry of a strip of a spatially tiled run (tiled.py): the tensors of resolution level l are widths[l] columns wide
// (own columns + that level's halo), and the 2x2 max-pool of a level-l tensor (widths[l] / 2 columns) lands at column
// pool_xoff[l] of the level-(l+1) tensor, whose outer columns only ever hold what the neighbour sends.  NULL: one image,
// widths[l] = W >> l, no offsets.
struct StripGeom {
    const int* widths;
    const int* pool_xoff;
};
static void layer_hw(int conv, int H, int W, const StripGeom& g, int* h, int* w) {
    layer_hw(conv, H, W, h, w);
    if (g.widths != nullptr) *w = g.widths[pools_before(conv)];
}

static int launch_conv_simt(int mode, bool pre, const float* X, const float* Wt, const float* bias, float* Y, const float* seed,
                       const float* mask, int H, int W, int Cin, int Cout, uint32_t* y_absmax, cudaStream_t st) {
    ADPST_REQUIRE(Cout % CT_N == 0, "conv3x3: Cout=%d must be a multiple of %d", Cout, CT_N);
    ADPST_REQUIRE(pre || Cin % CT_K == 0, "conv3x3: Cin=%d must be a multiple of %d", Cin, CT_K);
    const int tw = (W + CT_W - 1) / CT_W, th = (H + CT_H - 1) / CT_H;
    dim3 grid(tw * th, Cout / CT_N);
    if (mode == MODE_FWD && pre)
        conv3x3_simt_kernel<MODE_FWD, true><<<grid, CT_THREADS, 0, st>>>(X, Wt, bias, Y, seed, mask, H, W, Cin, Cout, tw, y_absmax);
    else if (mode == MODE_FWD)
        conv3x3_simt_kernel<MODE_FWD, false><<<grid, CT_THREADS, 0, st>>>(X, Wt, bias, Y, seed, mask, H, W, Cin, Cout, tw, y_absmax);
    else
        conv3x3_simt_kernel<MODE_BWD, false><<<grid, CT_THREADS, 0, st>>>(X, Wt, bias, Y, seed, mask, H, W, Cin, Cout, tw, y_absmax);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

// One convolution of the network: tensor-core path when the shape allows it, exact-fp32 CUDA-core path otherwise
// (block1_conv1: Cin = 3) or when the handle was switched to CONV_PATH_SIMT (validation).
// x_absmax: slot with max|X| (NULL: measured here with an extra pass over X); y_absmax (may be NULL): slot that receives
// max|Y| (must have been zeroed; every kernel family records it in its epilogue).
// pool_out (forward, may be NULL): where the following 2x2 max-pool goes; *pooled is set when the kernel wrote it.
static int launch_conv(adpst_vgg* h, int i, int mode, const float* X, float* Y, const float* seed, const float* mask,
                       int lh, int lw, const uint32_t* x_absmax, uint32_t* y_absmax, cudaStream_t st,
                       float* pool_out = nullptr, bool* pooled = nullptr, int pool_pitch = 0, int pool_xoff = 0) {
    const bool grad = (mode == MODE_BWD);
    const int K = grad ? conv_cout(i) : conv_cin(i), N = grad ? conv_cin(i) : conv_cout(i);
    if (h->conv_path == CONV_PATH_TENSOR && h->tc_ready && i > 0 && conv_tc_eligible(K, N)) {
        if (x_absmax == nullptr) {
            uint32_t* scratch = h->amax + AMAX_SCRATCH;
            int rc = launch_absmax(X, size_t(lh) * lw * K, scratch, st);
            if (rc != ADPST_OK) return rc;
            x_absmax = scratch;
        }
        if (pooled) *pooled = (pool_out != nullptr && !grad);
        return launch_conv_tc(h, i, grad ? 1 : 0, X, Y, seed, mask, lh, lw, K, N, x_absmax, y_absmax, grad ? nullptr : pool_out, st,
                              pool_pitch > 0 ? pool_pitch : lw / 2, pool_xoff);
    }
    return launch_conv_simt(mode, i == 0 && !grad, X, grad ? h->wb[i] : h->wf[i], grad ? nullptr : h->bias[i], Y, seed, mask,
                            lh, lw, K, N, y_absmax, st);
}

}  // namespace adpst

extern "C" {

int adpst_vgg_conv_shape(int i, int H, int W, int* h, int* w, int* c) {
    using namespace adpst;
    ADPST_REQUIRE(i >= 0 && i < kNumConv && h && w && c, "vgg_conv_shape: bad argument");
    layer_hw(i, H, W, h, w);
    *c = conv_cout(i);
    return ADPST_OK;
}

int adpst_vgg_pool_shape(int j, int H, int W, int* h, int* w, int* c) {
    using namespace adpst;
    ADPST_REQUIRE(j >= 0 && j < ADPST_VGG_NUM_POOL && h && w && c, "vgg_pool_shape: bad argument");
    *h = H >> (j + 1);
    *w = W >> (j + 1);
    *c = conv_cout(kPoolAfter[j]);
    return ADPST_OK;
}

int adpst_vgg_create(const float* const* kernels_dev, const float* const* biases_dev, adpst_stream_t stream,
                     adpst_vgg** out) {
    using namespace adpst;
    ADPST_REQUIRE(kernels_dev && biases_dev && out, "vgg_create: NULL argument");
    *out = nullptr;
    cudaStream_t st = as_stream(stream);
    auto* h = new adpst_vgg();
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < kNumConv && e == cudaSuccess; ++i) {
        if (!kernels_dev[i] || !biases_dev[i]) {
            adpst_vgg_destroy(h);
            return fail(ADPST_ERR_INVALID, "vgg_create: kernel/bias %d is NULL", i);
        }
        const size_t nw = size_t(9) * conv_cin(i) * conv_cout(i);
        e = cudaMalloc(reinterpret_cast<void**>(&h->wf[i]), nw * 4);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->bias[i]), size_t(conv_cout(i)) * 4);
        if (e == cudaSuccess && i > 0) e = cudaMalloc(reinterpret_cast<void**>(&h->wb[i]), nw * 4);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->wf[i], kernels_dev[i], nw * 4, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(h->bias[i], biases_dev[i], size_t(conv_cout(i)) * 4, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess && i > 0) {
            flip_transpose_weights_kernel<<<stream_grid(nw), 256, 0, st>>>(h->wf[i], h->wb[i], conv_cin(i), conv_cout(i));
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->wg0), 9 * 3 * 64 * 4);
    if (e == cudaSuccess) {
        conv1_image_weights_kernel<<<(9 * 3 * 64 + 255) / 256, 256, 0, st>>>(h->wf[0], h->wg0);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        adpst_vgg_destroy(h);
        return fail(ADPST_ERR_CUDA, "vgg_create: %s", cudaGetErrorString(e));
    }
    if (cudaMalloc(reinterpret_cast<void**>(&h->amax), AMAX_SLOTS * sizeof(uint32_t)) != cudaSuccess ||
        cudaMemsetAsync(h->amax, 0, AMAX_SLOTS * sizeof(uint32_t), st) != cudaSuccess) {
        adpst_vgg_destroy(h);
        return fail(ADPST_ERR_CUDA, "vgg_create: cannot allocate the scale slots");
    }
    for (int i = 1; i < kNumConv; ++i) {
        int rc = prepare_tc_weights(h, i, st);
        if (rc != ADPST_OK) { adpst_vgg_destroy(h); return rc; }
    }
    h->tc_ready = true;
    *out = h;
    return ADPST_OK;
}

void adpst_vgg_destroy(adpst_vgg* h) {
    if (!h) return;
    for (int i = 0; i < adpst::kNumConv; ++i) {
        if (h->wf[i]) cudaFree(h->wf[i]);
        if (h->wb[i]) cudaFree(h->wb[i]);
        if (h->bias[i]) cudaFree(h->bias[i]);
        for (int g = 0; g < 2; ++g) {
            if (h->tc_hi[g][i]) cudaFree(h->tc_hi[g][i]);
            if (h->tc_lo[g][i]) cudaFree(h->tc_lo[g][i]);
        }
    }
    if (h->wg0) cudaFree(h->wg0);
    if (h->amax) cudaFree(h->amax);
    delete h;
}

// convolutions first..last of the network; the input of conv `first` is the image (first == 0), the pooled tensor before
// it, or the previous convolution's output
static int vgg_forward_range(adpst_vgg* h, const float* image_dev, int H, int W, float* const* acts_dev, float* const* pools_dev,
                             int first, int last, cudaStream_t st, adpst::StripGeom geom = {nullptr, nullptr}) {
    using namespace adpst;
    const float* x = image_dev;
    if (first > 0) {
        x = acts_dev[first - 1];
        for (int j = 0; j < ADPST_VGG_NUM_POOL; ++j)
            if (kPoolAfter[j] == first - 1) x = pools_dev[j];
        ADPST_REQUIRE(x != nullptr, "vgg_forward: the input of conv %d is NULL", first);
    }
    ADPST_CUDA_CHECK(cudaMemsetAsync(h->amax + AMAX_ACT + first, 0, (last - first + 1) * sizeof(uint32_t), st));
    for (int i = first; i <= last; ++i) {
        int lh, lw;
        layer_hw(i, H, W, geom, &lh, &lw);
        ADPST_REQUIRE(acts_dev[i] != nullptr, "vgg_forward: acts[%d] is NULL", i);
        // a max-pool keeps the maximum of a post-ReLU map, so the pooled tensor shares the slot of the conv before it
        float* pool_out = nullptr;                 // the tensor-core kernel pools in its epilogue
        int pool_pitch = lw / 2, pool_xoff = 0;
        for (int j = 0; j < ADPST_VGG_NUM_POOL; ++j)
            if (kPoolAfter[j] == i && pools_dev[j] != nullptr) {
                pool_out = pools_dev[j];
                if (geom.widths != nullptr) {
                    pool_pitch = geom.widths[j + 1];
                    pool_xoff = geom.pool_xoff[j];
                    ADPST_REQUIRE(pool_xoff >= 0 && pool_xoff + lw / 2 <= pool_pitch, "vgg_forward: pool %d does not fit its strip", j);
                }
            }
        bool pooled = false;
        int rc = launch_conv(h, i, MODE_FWD, x, acts_dev[i], nullptr, nullptr, lh, lw, i > 0 ? h->amax + AMAX_ACT + i - 1 : nullptr,
                             h->amax + AMAX_ACT + i, st, pool_out, &pooled, pool_pitch, pool_xoff);
        if (rc != ADPST_OK) return rc;
        x = acts_dev[i];
        if (pool_out != nullptr) {
            if (!pooled) {
                const size_t items = size_t(lh / 2) * (lw / 2) * (conv_cout(i) / 4);
                if (items > 0) {
                    maxpool2_kernel<<<stream_grid(items), 256, 0, st>>>(acts_dev[i], pool_out, lh, lw, conv_cout(i), pool_pitch,
                                                                        pool_xoff);
                    ADPST_LAUNCH_CHECK();
                }
            }
            x = pool_out;
        }
    }
    return ADPST_OK;
}

int adpst_vgg_forward(adpst_vgg* h, const float* image_dev, int H, int W, float* const* acts_dev,
                      float* const* pools_dev, int last, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(h && image_dev && acts_dev && pools_dev, "vgg_forward: NULL argument");
    ADPST_REQUIRE(last >= 0 && last < kNumConv, "vgg_forward: last=%d out of range", last);
    ADPST_REQUIRE(H >= 1 && W >= 1, "vgg_forward: empty image");
    ADPST_REQUIRE((H >> pools_before(last)) >= 1 && (W >> pools_before(last)) >= 1,
                  "vgg_forward: %dx%d image is too small for conv %d", H, W, last);
    // the pool after conv `last` is not part of the extractor
    float* pools[ADPST_VGG_NUM_POOL];
    for (int j = 0; j < ADPST_VGG_NUM_POOL; ++j) {
        pools[j] = kPoolAfter[j] < last ? pools_dev[j] : nullptr;
        ADPST_REQUIRE(kPoolAfter[j] >= last || pools[j] != nullptr, "vgg_forward: pools[%d] is NULL", j);
    }
    return vgg_forward_range(h, image_dev, H, W, acts_dev, pools, 0, last, as_stream(stream));
}

static int check_strip_geom(const int* level_widths, const int* pool_col_offset, int W, int last, const char* who) {
    using namespace adpst;
    ADPST_REQUIRE((level_widths == nullptr) == (pool_col_offset == nullptr), "%s: level_widths and pool_col_offset go together", who);
    if (level_widths == nullptr) return ADPST_OK;
    ADPST_REQUIRE(level_widths[0] == W, "%s: level_widths[0]=%d is not the image width %d", who, level_widths[0], W);
    for (int l = 0; l <= pools_before(last); ++l) ADPST_REQUIRE(level_widths[l] >= 1, "%s: level %d is empty", who, l);
    for (int l = 0; l < pools_before(last); ++l)
        ADPST_REQUIRE(pool_col_offset[l] >= 0 && pool_col_offset[l] + level_widths[l] / 2 <= level_widths[l + 1],
                      "%s: the pool of level %d (%d columns at offset %d) does not fit level %d (%d columns)", who, l,
                      level_widths[l] / 2, pool_col_offset[l], l + 1, level_widths[l + 1]);
    return ADPST_OK;
}

int adpst_vgg_forward_range(adpst_vgg* h, const float* image_dev, int H, int W, float* const* acts_dev,
                            float* const* pools_dev, int first, int last, const int* level_widths, const int* pool_col_offset,
                            adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(h && acts_dev && pools_dev, "vgg_forward_range: NULL argument");
    ADPST_REQUIRE(first >= 0 && first <= last && last < kNumConv, "vgg_forward_range: bad range %d..%d", first, last);
    ADPST_REQUIRE(first > 0 || image_dev != nullptr, "vgg_forward_range: the image is NULL");
    ADPST_REQUIRE(H >= 1 && W >= 1 && (H >> pools_before(last)) >= 1 && (W >> pools_before(last)) >= 1,
                  "vgg_forward_range: %dx%d image is too small for conv %d", H, W, last);
    int rc = check_strip_geom(level_widths, pool_col_offset, W, last, "vgg_forward_range");
    if (rc != ADPST_OK) return rc;
    return vgg_forward_range(h, image_dev, H, W, acts_dev, pools_dev, first, last, as_stream(stream),
                             StripGeom{level_widths, pool_col_offset});
}

int adpst_absmax_update(const float* x_dev, size_t n, uint32_t* slot_dev, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(x_dev && slot_dev, "absmax_update: NULL argument");
    return launch_absmax(x_dev, n, slot_dev, as_stream(stream), false);
}

int adpst_vgg_set_conv_path(adpst_vgg* h, int path) {
    using namespace adpst;
    ADPST_REQUIRE(h && (path == CONV_PATH_TENSOR || path == CONV_PATH_SIMT), "vgg_set_conv_path: bad argument");
    h->conv_path = path;
    return ADPST_OK;
}

int adpst_absmax(const float* x_dev, size_t n, uint32_t* slot_dev, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(x_dev && slot_dev, "absmax: NULL argument");
    return launch_absmax(x_dev, n, slot_dev, as_stream(stream));
}

const uint32_t* adpst_vgg_act_absmax(const adpst_vgg* h, int i) {
    if (!h || i < 0 || i >= adpst::kNumConv || !h->amax) return nullptr;
    return h->amax + adpst::AMAX_ACT + i;
}

int adpst_vgg_conv_forward(adpst_vgg* h, int i, const float* x_dev, int lh, int lw, float* y_dev,
                           const uint32_t* x_absmax_dev, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(h && x_dev && y_dev && i >= 0 && i < kNumConv && lh > 0 && lw > 0, "vgg_conv_forward: bad argument");
    return launch_conv(h, i, MODE_FWD, x_dev, y_dev, nullptr, nullptr, lh, lw, x_absmax_dev, nullptr, as_stream(stream));
}

int adpst_vgg_conv_dgrad(adpst_vgg* h, int i, const float* dpre_dev, int lh, int lw, float* dx_dev,
                         const uint32_t* dpre_absmax_dev, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(h && dpre_dev && dx_dev && i >= 1 && i < kNumConv && lh > 0 && lw > 0, "vgg_conv_dgrad: bad argument");
    return launch_conv(h, i, MODE_BWD, dpre_dev, dx_dev, nullptr, nullptr, lh, lw, dpre_absmax_dev, nullptr, as_stream(stream));
}

// Backward through convs last..first.
// Entry: at the top (grad_in == NULL: the seed of conv `last` goes through its ReLU mask); below a pool (conv `last` is followed
// by a pool: grad_in = dLoss/d(pooled output of conv last), routed through the un-pool, the ReLU mask and the seed of conv last);
// or in the middle of a block (no pool after conv `last`: grad_in = dLoss/d(pre-activation of conv last), already masked).
// Exit: at the image (first == 0: out = dLoss/d(image)); above a pool (conv `first` follows a pool: out = dLoss/d(pooled input
// of conv first)); or in the middle of a block (out = dLoss/d(pre-activation of conv first - 1), i.e. with the seed and the
// ReLU mask of conv first - 1 applied -- exactly what the next call takes as grad_in).
static int vgg_backward_range(adpst_vgg* h, int H, int W, const float* const* acts_dev, const float* const* seeds_dev, int first,
                              int last, const float* grad_in, float* scratch0_dev, float* scratch1_dev, float* out_dev,
                              cudaStream_t st, adpst::StripGeom geom = {nullptr, nullptr}) {
    using namespace adpst;
    float* cur = scratch0_dev;   // holds dLoss/d(pre-activation of conv i)
    float* nxt = scratch1_dev;
    int lh, lw;
    layer_hw(last, H, W, geom, &lh, &lw);
    // where the gradient w.r.t. the pooled output of conv i lives: pitch and column offset (see StripGeom)
    auto pooled_geom = [&](int i, int w_i, int* pitch, int* xoff) {
        *pitch = w_i / 2;
        *xoff = 0;
        if (geom.widths != nullptr) {
            const int l = pools_before(i);
            *pitch = geom.widths[l + 1];
            *xoff = geom.pool_xoff[l];
        }
    };
    uint32_t* gmax = h->amax + AMAX_GRAD;          // gmax[i]: max|dLoss/d(pre-activation of conv i)|
    bool pooled_top = false;
    for (int j = 0; j < ADPST_VGG_NUM_POOL; ++j) pooled_top |= (kPoolAfter[j] == last);
    const bool mid_entry = grad_in != nullptr && !pooled_top;
    bool mid_exit = first > 0;
    for (int j = 0; j < ADPST_VGG_NUM_POOL; ++j)
        if (kPoolAfter[j] == first - 1) mid_exit = false;
    // slots written in this call: a mid-block entry arrives with gmax[last] already set by the producer of grad_in; a
    // mid-block exit also writes gmax[first - 1]
    const int zlo = first - (mid_exit ? 1 : 0), zhi = last - (mid_entry ? 1 : 0);
    if (zhi >= zlo) ADPST_CUDA_CHECK(cudaMemsetAsync(gmax + zlo, 0, (zhi - zlo + 1) * sizeof(uint32_t), st));
    if (grad_in == nullptr) {
        ADPST_REQUIRE(seeds_dev[last] != nullptr, "vgg_backward: the seed of the last layer (%d) is required", last);
        const size_t n4 = size_t(lh) * lw * conv_cout(last) / 4;
        relu_mask_kernel<<<stream_grid(n4), 256, 0, st>>>(acts_dev[last], seeds_dev[last], cur, n4, gmax + last);
        ADPST_LAUNCH_CHECK();
    } else if (pooled_top) {
        const size_t items = size_t((lh + 1) / 2) * ((lw + 1) / 2) * (conv_cout(last) / 4);
        int dpp, dpx;
        pooled_geom(last, lw, &dpp, &dpx);
        unpool_relu_kernel<<<stream_grid(items), 256, 0, st>>>(acts_dev[last], grad_in, seeds_dev[last], cur, lh, lw,
                                                               conv_cout(last), gmax + last, dpp, dpx);
        ADPST_LAUNCH_CHECK();
    } else {
        ADPST_CUDA_CHECK(cudaMemcpyAsync(cur, grad_in, size_t(lh) * lw * conv_cout(last) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    for (int i = last; i >= (first > 0 ? first : 1); --i) {
        layer_hw(i, H, W, geom, &lh, &lw);
        bool pooled_input = false;
        for (int j = 0; j < ADPST_VGG_NUM_POOL; ++j) pooled_input |= (kPoolAfter[j] == i - 1);
        if (!pooled_input) {
            // input of conv i is the post-ReLU output of conv i-1 at the same resolution
            float* dst = (i == first) ? out_dev : nxt;              // (first > 0 here: leave in the middle of a block)
            int rc = launch_conv(h, i, MODE_BWD, cur, dst, seeds_dev[i - 1], acts_dev[i - 1], lh, lw, gmax + i, gmax + i - 1, st);
            if (rc != ADPST_OK) return rc;
            if (i == first) return ADPST_OK;
            float* t = cur; cur = nxt; nxt = t;
        } else if (i == first) {
            // leave above the pool: gradient w.r.t. the pooled tensor straight into the caller's buffer
            return launch_conv(h, i, MODE_BWD, cur, out_dev, nullptr, nullptr, lh, lw, gmax + i, nullptr, st);
        } else {
            // gradient w.r.t. the pooled tensor, then route through the pool + ReLU of conv i-1
            int rc = launch_conv(h, i, MODE_BWD, cur, nxt, nullptr, nullptr, lh, lw, gmax + i, nullptr, st);
            if (rc != ADPST_OK) return rc;
            int ph, pw;
            layer_hw(i - 1, H, W, geom, &ph, &pw);
            const size_t items = size_t((ph + 1) / 2) * ((pw + 1) / 2) * (conv_cout(i - 1) / 4);
            int dpp, dpx;
            pooled_geom(i - 1, pw, &dpp, &dpx);
            unpool_relu_kernel<<<stream_grid(items), 256, 0, st>>>(acts_dev[i - 1], nxt, seeds_dev[i - 1], cur, ph, pw,
                                                                   conv_cout(i - 1), gmax + i - 1, dpp, dpx);
            ADPST_LAUNCH_CHECK();
        }
    }
    ADPST_ONCE_PER_DEVICE(ADPST_CUDA_CHECK(cudaFuncSetAttribute(conv1_dgrad_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IG_SMEM)));
    dim3 grid((W + IG_TW - 1) / IG_TW, (H + IG_TH - 1) / IG_TH);
    conv1_dgrad_image_kernel<<<grid, IG_THREADS, IG_SMEM, st>>>(cur, h->wg0, out_dev, H, W);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

int adpst_vgg_backward(adpst_vgg* h, int H, int W, const float* const* acts_dev, const float* const* pools_dev,
                       const float* const* seeds_dev, int last, float* scratch0_dev, float* scratch1_dev,
                       float* dimage_dev, adpst_stream_t stream) {
    using namespace adpst;
    (void)pools_dev;
    ADPST_REQUIRE(h && acts_dev && seeds_dev && scratch0_dev && scratch1_dev && dimage_dev, "vgg_backward: NULL argument");
    ADPST_REQUIRE(last >= 0 && last < kNumConv, "vgg_backward: last=%d out of range", last);
    return vgg_backward_range(h, H, W, acts_dev, seeds_dev, 0, last, nullptr, scratch0_dev, scratch1_dev, dimage_dev,
                              as_stream(stream));
}

int adpst_vgg_backward_range(adpst_vgg* h, int H, int W, const float* const* acts_dev, const float* const* seeds_dev, int first,
                             int last, const float* grad_in_dev, float* scratch0_dev, float* scratch1_dev, float* out_dev,
                             const int* level_widths, const int* pool_col_offset, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(h && acts_dev && seeds_dev && scratch0_dev && scratch1_dev && out_dev, "vgg_backward_range: NULL argument");
    ADPST_REQUIRE(first >= 0 && first <= last && last < kNumConv, "vgg_backward_range: bad range %d..%d", first, last);
    // (the gradient w.r.t. the pooled output of conv `last` lives one level up)
    int rc = check_strip_geom(level_widths, pool_col_offset, W, last + 1 < kNumConv ? last + 1 : last, "vgg_backward_range");
    if (rc != ADPST_OK) return rc;
    return vgg_backward_range(h, H, W, acts_dev, seeds_dev, first, last, grad_in_dev, scratch0_dev, scratch1_dev, out_dev,
                              as_stream(stream), StripGeom{level_widths, pool_col_offset});
}

const uint32_t* adpst_vgg_grad_absmax(const adpst_vgg* h, int i) {
    if (!h || i < 0 || i >= adpst::kNumConv || !h->amax) return nullptr;
    return h->amax + adpst::AMAX_GRAD + i;
}

}  // extern "C"
