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

// coef: [(y W + x)][12] (one 48-byte record per pixel), LI: (H,W,3), colq[x]: sum over the column of I . (L I) (pre-zeroed)
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
        const size_t o = size_t(y) * W + x;
        float r[DIA_P];
#pragma unroll
        for (int p = 0; p < DIA_P; ++p) {
            const int ty = y + dia_dy(p), tx = x + dia_dx(p);
            r[p] = (ty < H && tx >= 0 && tx < W) ? float(acc[p]) : 0.0f;
        }
        float4* rec = reinterpret_cast<float4*>(coef + o * DIA_P);
        rec[0] = make_float4(r[0], r[1], r[2], r[3]);
        rec[1] = make_float4(r[4], r[5], r[6], r[7]);
        rec[2] = make_float4(r[8], r[9], r[10], r[11]);
        const float* pi = img + o * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            LI[o * 3 + c] = float(li[c]);
            q = fma(double(pi[c]), li[c], q);
        }
    }
    sq[threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && x < W) atomicAdd(colq + x, sq[0][threadIdx.x] + sq[1][threadIdx.x] + sq[2][threadIdx.x] + sq[3][threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------------------
// mat-vec.  CTA = 32 x 16 output pixels, 256 threads, warp w owns tile rows w and w + 8.  d = x - I of the tile plus a
// 2-pixel halo is staged in shared memory as float4 (coalesced 4-byte loads of the interleaved RGB rows, no integer
// division in the loop).  Every pixel then reads its own 48-byte coefficient record (3 x LDG.128) and one coefficient from
// the record of each of its 12 backward neighbours (L1 hits: those records are the forward loads of the neighbouring
// threads), one LDS.128 per neighbour, 6 float32 operations per neighbour and channel.
// x^T L x: per pixel  x_i . (L d)_i + d_i . (L I)_i  (float32 dot of six terms, exact to 1e-7 of its largest term), summed
// over pixels in float64; the constant I^T L I comes from the per-column float64 sums of the build.
// ---------------------------------------------------------------------------------------------------------------
constexpr int DT_W = 32, DT_H = 16, DT_THREADS = 256, DT_PW = DT_W + 4, DT_PH = DT_H + 4;

__global__ void __launch_bounds__(DT_THREADS)
lap_dia_kernel(const float* __restrict__ x, const float* __restrict__ img, const float* __restrict__ LI,
               const float* __restrict__ coef, float* __restrict__ y, double* __restrict__ partial, int H, int W, float y_scale,
               int qlo, int qhi, unsigned int* __restrict__ ticket, double* __restrict__ xLx_out, const double* __restrict__ colq) {
    __shared__ float4 sd[DT_PH][DT_PW];
    __shared__ double sRed[32];
    __shared__ bool sLast;
    const int x0 = blockIdx.x * DT_W, y0 = blockIdx.y * DT_H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_elems = W * 3;
    // ---- stage d = x - I (zero outside the image: the coefficients that point there are zero as well)
    {
        int soff[4], ge[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = lane + 32 * k;                           // element of the 108-float tile row
            soff[k] = (e / 3) * 4 + (e % 3);
            ge[k] = (x0 - 2) * 3 + e;
            if (e >= DT_PW * 3 || ge[k] < 0 || ge[k] >= row_elems) ge[k] = -1;
        }
        float* sflat = reinterpret_cast<float*>(&sd[0][0]);
        for (int r = warp; r < DT_PH; r += DT_THREADS / 32) {
            const int gy = y0 - 2 + r;
            const bool row_ok = gy >= 0 && gy < H;
            const size_t g = size_t(row_ok ? gy : 0) * row_elems;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (lane + 32 * k < DT_PW * 3) {
                    float v = 0.f;
                    if (row_ok && ge[k] >= 0) v = __ldg(x + g + ge[k]) - __ldg(img + g + ge[k]);
                    sflat[r * (DT_PW * 4) + soff[k]] = v;
                }
            }
        }
    }
    __syncthreads();
    const int gx = x0 + lane;
    // which backward neighbours exist (their records hold the coefficient that points at this pixel)
    bool okx[5];
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) okx[dx + 2] = gx - dx >= 0 && gx - dx < W;
    const size_t W12 = size_t(W) * DIA_P;
    double qacc = 0.0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int ty = warp + half * 8, gy = y0 + ty;
        if (gx < W && gy < H) {
            const size_t o = size_t(gy) * W + gx;
            const float* cb = coef + o * DIA_P;
            const float4 c0 = __ldg(reinterpret_cast<const float4*>(cb));
            const float4 c1 = __ldg(reinterpret_cast<const float4*>(cb) + 1);
            const float4 c2 = __ldg(reinterpret_cast<const float4*>(cb) + 2);
            const float cf[DIA_P] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
            const float* cbr[3] = {cb, cb - W12, cb - 2 * W12};            // records of rows gy, gy - 1, gy - 2
            const bool oky[3] = {true, gy >= 1, gy >= 2};
            const float4 di = sd[ty + 2][lane + 2];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int p = 0; p < DIA_P; ++p) {
                const int dy = dia_dy(p), dx = dia_dx(p);
                {   // forward neighbour i + delta: own coefficient (stored as zero when the neighbour is outside the image)
                    const float4 dj = sd[ty + 2 + dy][lane + 2 + dx];
                    a0 = fmaf(cf[p], dj.x - di.x, a0); a1 = fmaf(cf[p], dj.y - di.y, a1); a2 = fmaf(cf[p], dj.z - di.z, a2);
                }
                if (oky[dy] && okx[dx + 2]) {   // backward neighbour i - delta: its coefficient for +delta (L_ij = L_ji)
                    const float cbk = __ldg(cbr[dy] - dx * DIA_P + p);
                    const float4 dj = sd[ty + 2 - dy][lane + 2 - dx];
                    a0 = fmaf(cbk, dj.x - di.x, a0); a1 = fmaf(cbk, dj.y - di.y, a1); a2 = fmaf(cbk, dj.z - di.z, a2);
                }
            }
            const float l0 = __ldg(LI + o * 3), l1 = __ldg(LI + o * 3 + 1), l2 = __ldg(LI + o * 3 + 2);
            if (partial != nullptr && gx >= qlo && gx < qhi) {
                const float x0v = __ldg(x + o * 3), x1v = __ldg(x + o * 3 + 1), x2v = __ldg(x + o * 3 + 2);
                float t = x0v * a0;
                t = fmaf(x1v, a1, t); t = fmaf(x2v, a2, t);
                t = fmaf(di.x, l0, t); t = fmaf(di.y, l1, t); t = fmaf(di.z, l2, t);
                qacc += double(t);
            }
            if (y != nullptr) {
                y[o * 3] = y_scale * (l0 + a0);
                y[o * 3 + 1] = y_scale * (l1 + a1);
                y[o * 3 + 2] = y_scale * (l2 + a2);
            }
        }
    }
    if (partial != nullptr) {
        // per-CTA partials summed in a fixed order by whichever CTA finishes last (deterministic, one launch, graph-replayable),
        // plus the constant I^T L I of the window from the per-column sums
        const int nblk = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
        const double tot = block_sum<double>(qacc, sRed);
        if (threadIdx.x == 0) {
            partial[blk] = tot;
            __threadfence();
            sLast = (atomicAdd(ticket, 1u) == unsigned(nblk - 1));
        }
        __syncthreads();
        if (sLast) {
            __threadfence();
            double a = 0.0;
            for (int i = threadIdx.x; i < nblk; i += blockDim.x) a += partial[i];
            for (int i = qlo + threadIdx.x; i < qhi; i += blockDim.x) a += colq[i];
            a = block_sum<double>(a, sRed);
            if (threadIdx.x == 0) { *xLx_out = a; *ticket = 0u; }
        }
    }
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
    unsigned int* ticket = reinterpret_cast<unsigned int*>(h->partials + h->npartials);
    lap_dia_kernel<<<grid, DT_THREADS, 0, st>>>(x, static_cast<const float*>(h->image), h->dia_LI, h->dia_coef, y,
                                                xLx ? h->partials : nullptr, h->H, h->W, float(y_scale), qlo, qhi, ticket, xLx,
                                                h->dia_qI);
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

}  // namespace adpst
