// elementwise.cu -- HBM-bound streaming kernels: Adam+clip, content MSE, mask resize, axpby.
//
// Replaces   style_transfer.py:321-326,342-343  (tf.optimizers.Adam.apply_gradients + clip_by_value)
//            components/loss.py:90-92           (content MSE and its gradient)
//            components/loss.py:112-113         (tf.image.resize of the masks)
#include "common.cuh"

namespace adpst {

static inline unsigned grid_for(size_t work_items, int threads, int per_sm = 8) {
    const size_t want = (work_items + threads - 1) / threads;
    const size_t cap = size_t(num_sms()) * per_sm;
    return unsigned(want < cap ? (want ? want : 1) : cap);
}

// ---------------------------------------------------------------------------------------------
// Adam (TF/Keras flavour) + clip to [0,1].  28 B per element: read g,m,v,x; write m,v,x.
// state[0] = completed steps t, state[1] = CTA ticket.  The last CTA to finish bumps t, so the launch
// can sit inside a CUDA graph and be replayed.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_one(float& x, float g, float& m, float& v, float b1, float b2, float alpha,
                                         float eps) {
    m = b1 * m + (1.0f - b1) * g;
    v = b2 * v + (1.0f - b2) * g * g;
    const float nx = x - alpha * m / (sqrtf(v) + eps);
    x = fminf(fmaxf(nx, 0.0f), 1.0f);
}

__global__ void __launch_bounds__(256)
adam_clip_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 size_t n, int32_t* __restrict__ state, float lr, float b1, float b2, float eps) {
    const int t = state[0] + 1;
    const float alpha = float(double(lr) * sqrt(1.0 - pow(double(b2), double(t))) / (1.0 - pow(double(b1), double(t))));
    const size_t n4 = n / 4;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    float4* x4 = reinterpret_cast<float4*>(x);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 xx = x4[i], mm = m4[i], vv = v4[i];
        const float4 gg = __ldg(g4 + i);
        adam_one(xx.x, gg.x, mm.x, vv.x, b1, b2, alpha, eps);
        adam_one(xx.y, gg.y, mm.y, vv.y, b1, b2, alpha, eps);
        adam_one(xx.z, gg.z, mm.z, vv.z, b1, b2, alpha, eps);
        adam_one(xx.w, gg.w, mm.w, vv.w, b1, b2, alpha, eps);
        x4[i] = xx; m4[i] = mm; v4[i] = vv;
    }
    if (blockIdx.x == 0) {
        for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) adam_one(x[i], g[i], m[i], v[i], b1, b2, alpha, eps);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(&state[1], 1);
        if (ticket == int(gridDim.x) - 1) { state[1] = 0; state[0] = t; }
    }
}

// ---------------------------------------------------------------------------------------------
// content layer: loss += scale * mean((t-o)^2); dOut (=|+=) scale * 2 (o-t) / n
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
content_kernel(const float* __restrict__ tgt, const float* __restrict__ out, size_t n, double loss_scale,
               double grad_scale, double* __restrict__ loss, float* __restrict__ dOut, int accumulate, double n_norm,
               int row_elems, int col_lo_elems, int col_hi_elems) {
    __shared__ double red[32];
    const float gs = float(2.0 * grad_scale / n_norm);
    double acc = 0.0;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = out[i] - tgt[i];
        // spatially tiled runs: the scalar counts only this rank's own columns, the gradient seed covers the halo too
        bool own = true;
        if (row_elems > 0) { const int e = int(i % size_t(row_elems)); own = e >= col_lo_elems && e < col_hi_elems; }
        if (own) acc += double(d) * double(d);
        if (dOut) dOut[i] = accumulate ? dOut[i] + gs * d : gs * d;
    }
    acc = block_sum<double>(acc, red);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, acc * loss_scale / n_norm);
}

// ---------------------------------------------------------------------------------------------
// bilinear resize, half-pixel centres, no antialias (tf.image.resize default == F.interpolate(bilinear,
// align_corners=False)).  Single channel.  Runs once per mask at set-up, not in the loop.
// ---------------------------------------------------------------------------------------------
__global__ void resize_bilinear_kernel(const float* __restrict__ src, int Hs, int Ws, float* __restrict__ dst, int Hd,
                                       int Wd) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= Hd || j >= Wd) return;
    src += size_t(blockIdx.z) * Hs * Ws;                 // plane of a batch
    dst += size_t(blockIdx.z) * Hd * Wd;
    const double sy = fmax((i + 0.5) * (double(Hs) / Hd) - 0.5, 0.0);
    const double sx = fmax((j + 0.5) * (double(Ws) / Wd) - 0.5, 0.0);
    const int y0 = min(int(sy), Hs - 1), x0 = min(int(sx), Ws - 1);
    const int y1 = min(y0 + 1, Hs - 1), x1 = min(x0 + 1, Ws - 1);
    const double ly = sy - y0, lx = sx - x0;
    const double v00 = src[size_t(y0) * Ws + x0], v01 = src[size_t(y0) * Ws + x1];
    const double v10 = src[size_t(y1) * Ws + x0], v11 = src[size_t(y1) * Ws + x1];
    const double top = v00 + (v01 - v00) * lx, bot = v10 + (v11 - v10) * lx;
    dst[size_t(i) * Wd + j] = float(top + (bot - top) * ly);
}

__global__ void __launch_bounds__(256)
axpby_kernel(float* __restrict__ out, const float* __restrict__ a, float alpha, const float* __restrict__ b, float beta,
             size_t n) {
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = alpha * a[i] + (b ? beta * b[i] : 0.0f);
}

__global__ void loss_finalize_kernel(const double* __restrict__ acc, double wc, double ws, double wp, double wtv,
                                     float* __restrict__ out) {
    const double c = acc[0], s = acc[1], p = acc[2], tv = acc[3];
    out[0] = float(c); out[1] = float(s); out[2] = 0.0f; out[3] = float(p); out[5] = float(tv);
    // loss.py:72: the terms are float32 tensors when they are weighted and added, in insertion order
    float total = float(wc) * float(c) + float(ws) * float(s) + (wp > 0.0 ? float(wp) * float(p) : 0.0f);
    if (wtv > 0.0) total += float(wtv) * float(tv);      // extension term, after the reference's
    out[4] = total;
}

// ---------------------------------------------------------------------------------------------
// Total variation (EXTENSION: not in the reference; semantics of tf.image.total_variation):
//   TV(x) = sum |x[y+1,x,c] - x[y,x,c]| + sum |x[y,x+1,c] - x[y,x,c]|          (anisotropic L1, no normalisation)
//   dTV/dx[y,x,c] = sgn(x[y,x]-x[y-1,x]) - sgn(x[y+1,x]-x[y,x]) + sgn(x[y,x]-x[y,x-1]) - sgn(x[y,x+1]-x[y,x]),  sgn(0) = 0
// One pass: 12 B/px read (the four neighbours come from L1/L2), 12 B/px written (24 when accumulating into an existing
// gradient buffer).  The value is reduced with warp shuffles, one float64 atomic per CTA.
// Every element owns its "down" and "right" differences; the scalar counts the elements of columns [col_lo, col_hi).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sgnf(float d) { return float(d > 0.f) - float(d < 0.f); }

__global__ void __launch_bounds__(256)
tv_kernel(const float* __restrict__ x, int H, int W, double loss_scale, float grad_scale, double* __restrict__ loss,
          float* __restrict__ dX, int accumulate, int col_lo, int col_hi) {
    __shared__ double red[32];
    const int row = W * 3;
    const size_t n = size_t(H) * row;
    double acc = 0.0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const int y = int(i / size_t(row)), e = int(i - size_t(y) * row), col = e / 3;
        const float v = __ldg(x + i);
        float g = 0.f;
        double a = 0.0;
        if (y > 0) g += sgnf(v - __ldg(x + i - row));
        if (col > 0) g += sgnf(v - __ldg(x + i - 3));
        if (y + 1 < H) { const float u = __ldg(x + i + row); g -= sgnf(u - v); a += fabs(double(u) - double(v)); }
        if (col + 1 < W) { const float u = __ldg(x + i + 3); g -= sgnf(u - v); a += fabs(double(u) - double(v)); }
        if (col >= col_lo && col < col_hi) acc += a;
        if (dX) dX[i] = accumulate ? fmaf(grad_scale, g, dX[i]) : grad_scale * g;
    }
    acc = block_sum<double>(acc, red);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, acc * loss_scale);
}

}  // namespace adpst

extern "C" {

int adpst_adam_clip_step(float* x_dev, const float* grad_dev, float* m_dev, float* v_dev, size_t n, int32_t* state_dev,
                         float lr, float beta1, float beta2, float epsilon, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(x_dev && grad_dev && m_dev && v_dev && state_dev, "adam_clip_step: NULL argument");
    ADPST_REQUIRE((reinterpret_cast<uintptr_t>(x_dev) | reinterpret_cast<uintptr_t>(grad_dev) |
                   reinterpret_cast<uintptr_t>(m_dev) | reinterpret_cast<uintptr_t>(v_dev)) % 16 == 0,
                  "adam_clip_step: buffers must be 16-byte aligned");
    if (n == 0) return ADPST_OK;
    adam_clip_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, as_stream(stream)>>>(x_dev, grad_dev, m_dev, v_dev, n, state_dev,
                                                                              lr, beta1, beta2, epsilon);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

int adpst_content_layer(const float* target_dev, const float* output_dev, size_t n, double loss_scale, double grad_scale,
                        double* loss_dev, float* dOut_dev, int accumulate, double n_norm, int w, int C, int col_lo, int col_hi,
                        adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(target_dev && output_dev && n > 0, "content_layer: NULL or empty input");
    ADPST_REQUIRE(w == 0 || (C > 0 && col_lo >= 0 && col_hi <= w && col_lo <= col_hi && n % (size_t(w) * C) == 0),
                  "content_layer: bad column window");
    if (n_norm <= 0.0) n_norm = double(n);
    content_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(target_dev, output_dev, n, loss_scale, grad_scale,
                                                                    loss_dev, dOut_dev, accumulate, n_norm, w * C, col_lo * C,
                                                                    col_hi * C);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

int adpst_resize_bilinear_batch(const float* src_dev, int n, int Hs, int Ws, float* dst_dev, int Hd, int Wd,
                                adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(src_dev && dst_dev && n > 0 && n <= 65535 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0, "resize_bilinear: bad argument");
    dim3 block(32, 8), grid((Wd + 31) / 32, (Hd + 7) / 8, n);
    resize_bilinear_kernel<<<grid, block, 0, as_stream(stream)>>>(src_dev, Hs, Ws, dst_dev, Hd, Wd);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

int adpst_resize_bilinear(const float* src_dev, int Hs, int Ws, float* dst_dev, int Hd, int Wd, adpst_stream_t stream) {
    return adpst_resize_bilinear_batch(src_dev, 1, Hs, Ws, dst_dev, Hd, Wd, stream);
}

int adpst_tv_loss(const float* x_dev, int H, int W, double loss_scale, double grad_scale, double* loss_dev, float* dX_dev,
                  int accumulate, int col_lo, int col_hi, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(x_dev && H > 0 && W > 0, "tv_loss: NULL or empty image");
    ADPST_REQUIRE(loss_dev || dX_dev, "tv_loss: nothing to compute (loss and dX NULL)");
    if (col_lo == 0 && col_hi == 0) col_hi = W;
    ADPST_REQUIRE(col_lo >= 0 && col_hi <= W && col_lo <= col_hi, "tv_loss: bad column window");
    tv_kernel<<<grid_for(size_t(H) * W * 3, 256), 256, 0, as_stream(stream)>>>(x_dev, H, W, loss_scale, float(grad_scale),
                                                                              loss_dev, dX_dev, accumulate, col_lo, col_hi);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

int adpst_loss_finalize(const double* acc_dev, double w_content, double w_style, double w_photo, double w_tv,
                        float* out_dev, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(acc_dev && out_dev, "loss_finalize: NULL argument");
    loss_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(acc_dev, w_content, w_style, w_photo, w_tv, out_dev);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

int adpst_axpby(float* out_dev, const float* a_dev, float alpha, const float* b_dev, float beta, size_t n,
                adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(out_dev && a_dev, "axpby: NULL argument");
    if (n == 0) return ADPST_OK;
    axpby_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(out_dev, a_dev, alpha, b_dev, beta, n);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

}  // extern "C"
