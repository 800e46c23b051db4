// laplacian_dia.cu -- the matting Laplacian as a precomputed 5x5 variable-coefficient stencil ("diagonal format").
//
// Replaces, for window radius 1 and float32 storage (the path components/loss.py:157-161 runs every iteration):
//   components/matting_v2.py:147-176   _matmul          (matrix-free in the reference)
//   components/matting_v3.py:50-51     _matmul          (tf.sparse.sparse_dense_matmul on the explicit COO matrix)
//   components/matting_v3.py:61-102    compute_laplacian (the explicit matrix itself; here built on the GPU, in float64)
//
// Why.  The reference evaluates this term in float64 (loss.py:160) and it has to: with x ~ I (the first iterations) L x is a
// ~1e-7 remainder of O(1) terms, and the window covariances are ill-conditioned (eps = 1e-7).  The matrix-free float64 kernel
// (laplacian.cu) therefore spends ~235 float64 operations per pixel in EVERY iteration and is bound by the float64 pipe at
// <= 0.4 of the HBM roofline.  But everything that needs float64 depends on the guide image I only:
//   * L is a symmetric (H W x H W) matrix with a 5x5 footprint (25 non-zeros per row) and zero row sums, for v2 (symmetric
//     padding folds the out-of-image part of the footprint back inside, still 5x5 and still symmetric: verified against the
//     oracle to 1e-14) as well as v3;
//   * its entries are O(1) and WELL conditioned as numbers: rounding them to float32 perturbs L x by 6e-8 |L| |x|.
// So the 12 "forward" coefficients per pixel (symmetry gives the other 12, the zero row sum the diagonal) and the vector
// L I are computed ONCE per image in float64, stored as float32, and every iteration evaluates
//     y_i = (L I)_i + sum_{j in 5x5, j != i} L_ij ((x_j - I_j) - (x_i - I_i))                 (float32 FMAs)
//     x^T L x = I^T L I (float64 constant) + sum_i [ x_i (L d)_i + d_i (L I)_i ],  d = x - I   (float64 accumulation)
// which is exact at x = I and within 2e-7 of max|y| otherwise (measured against the float64 oracle on uniform, smooth and
// grey images).  Per pixel the kernel moves 96 bytes (12 coefficients, x, I, L I in; y out) against 36 B/px algorithmic, and
// is HBM-bound: the 64 extra bytes buy back the ~235 float64 operations.
#include "laplacian.cuh"

namespace adpst {

constexpr int DIA_P = 12;
// forward half of the 5x5 neighbourhood: p = 0,1 -> (0,1),(0,2); 2..6 -> (1,-2..2); 7..11 -> (2,-2..2)
__host__ __device__ constexpr int dia_dy(int p) { return p < 2 ? 0 : (p < 7 ? 1 : 2); }
__host__ __device__ constexpr int dia_dx(int p) { return p < 2 ? p + 1 : (p < 7 ? p - 4 : p - 9); }
__host__ __device__ constexpr bool dia_forward(int dy, int dx) { return dy > 0 || (dy == 0 && dx > 0); }
__host__ __device__ constexpr int dia_index(int dy, int dx) { return dy == 0 ? dx - 1 : (dy == 1 ? dx + 4 : dx + 9); }

// v = M^-1 a for the symmetric positive definite 3x3 M = [m0 m1 m2; m1 m3 m4; m2 m4 m5], by Cholesky factorisation.
// Cofactor inversion is NOT good enough here: on windows whose colours lie on a line (grey images, two-colour edges) M has
// eigenvalues (s, eps, eps) with s/eps up to 1e6, and the determinant / cofactors cancel that many digits (measured: 1e-5
// error of the stencil coefficients on a grey image).  Cholesky is backward stable, like the LU of np.linalg.inv that the
// reference calls (matting_v2.py:52, matting_v3.py:92).
__device__ __forceinline__ void spd3_solve(const double m[6], double a0, double a1, double a2, double& v0, double& v1, double& v2) {
    const double l00 = sqrt(m[0]);
    const double i00 = 1.0 / l00;
    const double l10 = m[1] * i00, l20 = m[2] * i00;
    const double l11 = sqrt(m[3] - l10 * l10);
    const double i11 = 1.0 / l11;
    const double l21 = (m[4] - l20 * l10) * i11;
    const double l22 = sqrt(m[5] - l20 * l20 - l21 * l21);
    const double i22 = 1.0 / l22;
    const double z0 = a0 * i00;
    const double z1 = (a1 - l10 * z0) * i11;
    const double z2 = (a2 - l20 * z0 - l21 * z1) * i22;
    v2 = z2 * i22;
    v1 = (z1 - l21 * v2) * i11;
    v0 = (z0 - l10 * v1 - l20 * v2) * i00;
}

// ---------------------------------------------------------------------------------------------------------------
// build: one thread per pixel i, float64 throughout.
//   L_ij (i != j) = sum over windows k containing i and j' of  -(1/n) - c_i^T M_k^-1 c_j',   M_k = sum c c^T + eps Id,
//   c = I - mu_k,  summed over all j' (coordinates of the padded image) that the symmetric reflection maps onto j.
//   (L I)_i = sum_j L_ij (I_j - I_i)  is accumulated in the same loop from the float64 terms, and  I_i . (L I)_i  is added to a
//   per-column float64 sum (the constant part of x^T L x for any column window).
// v2: a window is centred on every pixel of the image extended by one reflected ring; v3: only windows that lie inside.
// DYN = false: i is at least two pixels away from every border, no reflection can occur, offsets are compile-time constants
// and the accumulators stay in registers.  DYN = true (the two-pixel border ring of v2): offsets after reflection are run-time.
// ---------------------------------------------------------------------------------------------------------------
template <bool V2, bool DYN>
__device__ __forceinline__ void dia_row(const float* __restrict__ img, int H, int W, int y, int x, double eps, double (&acc)[DIA_P],
                                        double (&li)[3]) {
    const float* pi = img + (size_t(y) * W + x) * 3;
    const double i0 = pi[0], i1 = pi[1], i2 = pi[2];
#pragma unroll(DYN ? 1 : 3)
    for (int wy = -1; wy <= 1; ++wy) {
#pragma unroll(DYN ? 1 : 3)
        for (int wx = -1; wx <= 1; ++wx) {
            const int cy = y + wy, cx = x + wx;
            if (!V2 && (cy < 1 || cy >= H - 1 || cx < 1 || cx >= W - 1)) continue;
            int ty[3], tx[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                ty[a] = (V2 && DYN) ? reflect_symmetric(cy - 1 + a, H) : cy - 1 + a;
                tx[a] = (V2 && DYN) ? reflect_symmetric(cx - 1 + a, W) : cx - 1 + a;
            }
            double c[9][3];
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int py = 0; py < 3; ++py)
#pragma unroll
                for (int px = 0; px < 3; ++px) {
                    const float* q = img + (size_t(ty[py]) * W + tx[px]) * 3;
                    c[py * 3 + px][0] = q[0]; c[py * 3 + px][1] = q[1]; c[py * 3 + px][2] = q[2];
                    s0 += c[py * 3 + px][0]; s1 += c[py * 3 + px][1]; s2 += c[py * 3 + px][2];
                }
            const double m0 = s0 * (1.0 / 9.0), m1 = s1 * (1.0 / 9.0), m2 = s2 * (1.0 / 9.0);
            double M[6] = {eps, 0.0, 0.0, eps, 0.0, eps};
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                c[k][0] -= m0; c[k][1] -= m1; c[k][2] -= m2;
                M[0] = fma(c[k][0], c[k][0], M[0]); M[1] = fma(c[k][0], c[k][1], M[1]); M[2] = fma(c[k][0], c[k][2], M[2]);
                M[3] = fma(c[k][1], c[k][1], M[3]); M[4] = fma(c[k][1], c[k][2], M[4]); M[5] = fma(c[k][2], c[k][2], M[5]);
            }
            const double a0 = i0 - m0, a1 = i1 - m1, a2 = i2 - m2;         // c_i
            double v0, v1, v2;
            spd3_solve(M, a0, a1, a2, v0, v1, v2);
#pragma unroll
            for (int py = 0; py < 3; ++py)
#pragma unroll
                for (int px = 0; px < 3; ++px) {
                    const int k = py * 3 + px;
                    const double val = -(1.0 / 9.0) - (v0 * c[k][0] + v1 * c[k][1] + v2 * c[k][2]);
                    // (L I)_i += val * (I_j' - I_i);  I_j' - I_i = c_j' - c_i (the window mean cancels); zero for j' = i
                    li[0] = fma(val, c[k][0] - a0, li[0]); li[1] = fma(val, c[k][1] - a1, li[1]); li[2] = fma(val, c[k][2] - a2, li[2]);
                    if (DYN) {
                        const int ddy = ty[py] - y, ddx = tx[px] - x;
                        if (dia_forward(ddy, ddx)) acc[dia_index(ddy, ddx)] += val;
                    } else {
                        const int ddy = wy + py - 1, ddx = wx + px - 1;          // compile-time after unrolling
                        if (dia_forward(ddy, ddx)) acc[dia_index(ddy, ddx)] += val;
                    }
                }
        }
    }
}

// coef: [12][H][W] planes, LI: [3][H][W] planes, colq[x]: sum over the column of I . (L I) (pre-zeroed)
template <bool V2>
__global__ void __launch_bounds__(128)
lap_dia_build_kernel(const float* __restrict__ img, float* __restrict__ coef, float* __restrict__ LI, double* __restrict__ colq,
                     int H, int W, double eps) {
    __shared__ double sq[4][32];
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y;
    double q = 0.0;
    if (x < W && y < H) {
        double acc[DIA_P], li[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int p = 0; p < DIA_P; ++p) acc[p] = 0.0;
        const bool border = V2 && (y < 2 || x < 2 || y >= H - 2 || x >= W - 2);
        if (border) {
            double accd[DIA_P];             // indexed at run time (local memory); kept apart so that `acc` stays in registers
#pragma unroll
            for (int p = 0; p < DIA_P; ++p) accd[p] = 0.0;
            dia_row<V2, true>(img, H, W, y, x, eps, accd, li);
#pragma unroll
            for (int p = 0; p < DIA_P; ++p) acc[p] = accd[p];
        } else {
            dia_row<V2, false>(img, H, W, y, x, eps, acc, li);
        }
        const size_t HW = size_t(H) * W, o = size_t(y) * W + x;
#pragma unroll
        for (int p = 0; p < DIA_P; ++p) {
            const int ty = y + dia_dy(p), tx = x + dia_dx(p);
            coef[p * HW + o] = (ty < H && tx >= 0 && tx < W) ? float(acc[p]) : 0.0f;
        }
        const float* pi = img + o * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            LI[c * HW + o] = float(li[c]);
            q = fma(double(pi[c]), li[c], q);
        }
    }
    sq[threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && x < W) atomicAdd(colq + x, sq[0][threadIdx.x] + sq[1][threadIdx.x] + sq[2][threadIdx.x] + sq[3][threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------------------
// mat-vec.  The kernel is bound by the L1 / shared-memory data pipe (one 128-byte wavefront per clock and SM) unless the
// per-pixel wavefront count is kept near the ~130 that 96 B/px of HBM traffic allow, so:
//   * CTA = 32 x 16 output pixels, 256 threads; a thread owns TWO vertically adjacent pixels, whose 5x5 neighbourhoods
//     share 4 of 6 rows: 30 LDS.128 for two pixels instead of 50;
//   * coefficients and L I are stored as planes ([12][H][W], [3][H][W]): a warp's load of one coefficient of 32 neighbouring
//     pixels is one 128-byte line, for the own ("forward") coefficients as well as for the ones read from the backward
//     neighbours' planes (L_ij = L_ji; those are L1 / L2 hits: the same lines are the forward loads of neighbouring warps);
//   * d = x - I of the tile plus a 2-pixel halo is staged in shared memory as float4 (coalesced 4-byte loads of the
//     interleaved RGB rows, no integer division in the loop);
//   * everything that does not depend on shared memory (the staged x / I elements, both pixels' forward coefficients, L I)
//     is requested before the first use, so every thread has ~50 independent loads in flight.
// x^T L x: per pixel  x_i . (L d)_i + d_i . (L I)_i  (float32 dot of six terms), summed over pixels in float64; the constant
// I^T L I comes from the per-column float64 sums of the build (lap_dia_sum_kernel).
// ---------------------------------------------------------------------------------------------------------------
constexpr int DT_W = 32, DT_H = 16, DT_THREADS = 256, DT_PW = DT_W + 4, DT_PH = DT_H + 4;

__global__ void __launch_bounds__(DT_THREADS, 3)
lap_dia_kernel(const float* __restrict__ x, const float* __restrict__ img, const float* __restrict__ LI,
               const float* __restrict__ coef, float* __restrict__ y, double* __restrict__ partial, int H, int W, float y_scale,
               int qlo, int qhi) {
    __shared__ float4 sd[DT_PH][DT_PW];
    __shared__ double sRed[32];
    const int x0 = blockIdx.x * DT_W, y0 = blockIdx.y * DT_H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_elems = W * 3;
    const int gx = x0 + lane;
    const size_t HW = size_t(H) * W;
    const bool want_q = partial != nullptr;
    // ---- (1) requests for the staged rows (3 tile rows x 4 elements per thread)
    int soff[4], ge[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = lane + 32 * k;                               // element of the 108-float tile row
        soff[k] = (e / 3) * 4 + (e % 3);
        ge[k] = (x0 - 2) * 3 + e;
        if (e >= DT_PW * 3 || ge[k] < 0 || ge[k] >= row_elems) ge[k] = -1;
    }
    float sx[3][4], si[3][4];
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int r = warp + it * (DT_THREADS / 32);
        const int gyr = y0 - 2 + r;
        const bool row_ok = r < DT_PH && gyr >= 0 && gyr < H;
        const size_t g = size_t(row_ok ? gyr : 0) * row_elems;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool ok = row_ok && ge[k] >= 0;
            sx[it][k] = ok ? __ldg(x + g + ge[k]) : 0.f;
            si[it][k] = ok ? __ldg(img + g + ge[k]) : 0.f;
        }
    }
    // ---- (2) requests for this thread's two pixels (rows ty, ty + 1): forward coefficients and L I
    const int ty = 2 * warp, gy = y0 + ty;
    bool live[2];
    float cf[2][DIA_P], lv[2][3];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        live[k] = gx < W && gy + k < H;
        const size_t o = live[k] ? size_t(gy + k) * W + gx : 0;
#pragma unroll
        for (int p = 0; p < DIA_P; ++p) cf[k][p] = live[k] ? __ldg(coef + p * HW + o) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) lv[k][c] = live[k] ? __ldg(LI + c * HW + o) : 0.f;
    }
    // ---- (3) d = x - I into shared memory (zero outside the image: the coefficients that point there are zero as well)
    {
        float* sflat = reinterpret_cast<float*>(&sd[0][0]);
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int r = warp + it * (DT_THREADS / 32);
            if (r < DT_PH) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (lane + 32 * k < DT_PW * 3) sflat[r * (DT_PW * 4) + soff[k]] = sx[it][k] - si[it][k];
            }
        }
    }
    __syncthreads();
    double qacc = 0.0;
    if (live[0]) {                                                  // (live[1] implies live[0])
        // which backward neighbours exist (their planes hold the coefficient that points at this pixel)
        bool okx[5];
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) okx[dx + 2] = gx + dx >= 0 && gx + dx < W;
        const float* cbase = coef + size_t(gy) * W + gx;             // plane 0 at pixel 0 of this thread
        const float4 dc0 = sd[ty + 2][lane + 2], dc1 = sd[ty + 3][lane + 2];
        float a[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
        for (int rr = 0; rr < 6; ++rr) {                             // tile rows ty + rr = image rows gy - 2 + rr
            float4 dr[5];
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) dr[dx + 2] = sd[ty + rr][lane + 2 + dx];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int dy = rr - 2 - k;                           // offset of this row from pixel k
                if (dy < -2 || dy > 2) continue;
                const float4 dc = k == 0 ? dc0 : dc1;
#pragma unroll
                for (int dx = -2; dx <= 2; ++dx) {
                    if (dy == 0 && dx == 0) continue;
                    float c;
                    if (dia_forward(dy, dx)) {
                        c = cf[k][dia_index(dy, dx)];                // own coefficient (stored as zero if the neighbour is outside)
                    } else {
                        // backward neighbour j = i + (dy, dx): its coefficient for the offset (-dy, -dx) that points back at i
                        const bool ok = okx[dx + 2] && gy + k + dy >= 0 && (k == 0 || live[1]);
                        c = ok ? __ldg(cbase + dia_index(-dy, -dx) * HW + (ptrdiff_t(k + dy) * W + dx)) : 0.f;
                    }
                    const float4 dj = dr[dx + 2];
                    a[k][0] = fmaf(c, dj.x - dc.x, a[k][0]); a[k][1] = fmaf(c, dj.y - dc.y, a[k][1]); a[k][2] = fmaf(c, dj.z - dc.z, a[k][2]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!live[k]) continue;
            const size_t o = size_t(gy + k) * W + gx;
            const float4 dc = k == 0 ? dc0 : dc1;
            if (want_q && gx >= qlo && gx < qhi) {
                // x_i: an L1 hit (this CTA staged the row a moment ago)
                float t = __ldg(x + o * 3) * a[k][0];
                t = fmaf(__ldg(x + o * 3 + 1), a[k][1], t); t = fmaf(__ldg(x + o * 3 + 2), a[k][2], t);
                t = fmaf(dc.x, lv[k][0], t); t = fmaf(dc.y, lv[k][1], t); t = fmaf(dc.z, lv[k][2], t);
                qacc += double(t);
            }
            if (y != nullptr) {
                y[o * 3] = y_scale * (lv[k][0] + a[k][0]);
                y[o * 3 + 1] = y_scale * (lv[k][1] + a[k][1]);
                y[o * 3 + 2] = y_scale * (lv[k][2] + a[k][2]);
            }
        }
    }
    if (want_q) {       // per-CTA partial; lap_dia_sum_kernel adds them up in a fixed order (deterministic, no atomics, no fence)
        const double tot = block_sum<double>(qacc, sRed);
        if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
}

// x^T L x = sum of the per-CTA partials + the constant I^T L I of the column window (per-column float64 sums of the build)
__global__ void __launch_bounds__(256)
lap_dia_sum_kernel(const double* __restrict__ partial, int n, const double* __restrict__ colq, int qlo, int qhi,
                   double* __restrict__ out) {
    __shared__ double red[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += partial[i];
    for (int i = qlo + threadIdx.x; i < qhi; i += blockDim.x) a += colq[i];
    a = block_sum<double>(a, red);
    if (threadIdx.x == 0) *out = a;
}

// ---------------------------------------------------------------------------------------------------------------
bool dia_eligible(const adpst_laplacian* h) { return h->R == 1 && h->io_dtype == ADPST_F32; }

void dia_free(adpst_laplacian* h) {
    device_free(h->dia_coef, h->stream);
    device_free(h->dia_LI, h->stream);
    device_free(h->dia_qI, h->stream);
    h->dia_coef = nullptr; h->dia_LI = nullptr; h->dia_qI = nullptr; h->dia_ready = false;
}

int dia_build(adpst_laplacian* h, cudaStream_t st) {
    const size_t HW = size_t(h->H) * h->W;
    int rc0 = ADPST_OK;
    if (!h->dia_coef) rc0 = device_alloc(reinterpret_cast<void**>(&h->dia_coef), HW * DIA_P * sizeof(float), st);
    if (rc0 == ADPST_OK && !h->dia_LI) rc0 = device_alloc(reinterpret_cast<void**>(&h->dia_LI), HW * 3 * sizeof(float), st);
    if (rc0 == ADPST_OK && !h->dia_qI) rc0 = device_alloc(reinterpret_cast<void**>(&h->dia_qI), sizeof(double) * h->W, st);
    if (rc0 != ADPST_OK) return rc0;
    ADPST_CUDA_CHECK(cudaMemsetAsync(h->dia_qI, 0, sizeof(double) * h->W, st));
    const float* img = static_cast<const float*>(h->image);
    dim3 block(32, 4), grid((h->W + 31) / 32, (h->H + 3) / 4);
    if (h->mode == ADPST_LAP_V2) lap_dia_build_kernel<true><<<grid, block, 0, st>>>(img, h->dia_coef, h->dia_LI, h->dia_qI, h->H, h->W, h->eps);
    else lap_dia_build_kernel<false><<<grid, block, 0, st>>>(img, h->dia_coef, h->dia_LI, h->dia_qI, h->H, h->W, h->eps);
    ADPST_LAUNCH_CHECK();
    h->dia_ready = true;
    return ADPST_OK;
}

int dia_matvec(adpst_laplacian* h, const float* x, float* y, double y_scale, double* xLx, cudaStream_t st) {
    if (!h->dia_ready) return fail(ADPST_ERR_INVALID, "laplacian: the diagonal-format operator has not been built");
    const int qlo = h->q_col_hi > h->q_col_lo ? h->q_col_lo : 0, qhi = h->q_col_hi > h->q_col_lo ? h->q_col_hi : h->W;
    dim3 grid((h->W + DT_W - 1) / DT_W, (h->H + DT_H - 1) / DT_H);
    if (int(grid.x * grid.y) > h->npartials)
        return fail(ADPST_ERR_INVALID, "laplacian: partial buffer too small (%d > %d)", int(grid.x * grid.y), h->npartials);
    lap_dia_kernel<<<grid, DT_THREADS, 0, st>>>(x, static_cast<const float*>(h->image), h->dia_LI, h->dia_coef, y,
                                                xLx ? h->partials : nullptr, h->H, h->W, float(y_scale), qlo, qhi);
    ADPST_LAUNCH_CHECK();
    if (xLx) {
        lap_dia_sum_kernel<<<1, 256, 0, st>>>(h->partials, int(grid.x * grid.y), h->dia_qI, qlo, qhi, xLx);
        ADPST_LAUNCH_CHECK();
    }
    return ADPST_OK;
}

}  // namespace adpst
