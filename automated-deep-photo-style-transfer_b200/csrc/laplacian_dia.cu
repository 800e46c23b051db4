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

__device__ __forceinline__ void sym3_inverse_f64(const double m[6], double inv[6]) {
    const double c00 = m[3] * m[5] - m[4] * m[4];
    const double c01 = m[2] * m[4] - m[1] * m[5];
    const double c02 = m[1] * m[4] - m[2] * m[3];
    const double c11 = m[0] * m[5] - m[2] * m[2];
    const double c12 = m[1] * m[2] - m[0] * m[4];
    const double c22 = m[0] * m[3] - m[1] * m[1];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double id = 1.0 / det;
    inv[0] = c00 * id; inv[1] = c01 * id; inv[2] = c02 * id;
    inv[3] = c11 * id; inv[4] = c12 * id; inv[5] = c22 * id;
}

// ---------------------------------------------------------------------------------------------------------------
// build: one thread per pixel i computes the forward part of row i of L in float64.
//   L_ij (i != j) = sum over windows k containing i and j' of  -(1/n) - c_i^T M_k^-1 c_j',   M_k = sum c c^T + eps Id,
//   c = I - mu_k,  summed over all j' (coordinates of the padded image) that the symmetric reflection maps onto j.
// v2: a window is centred on every pixel of the image extended by one reflected ring; v3: only windows that lie inside.
// DYN = false: i is at least two pixels away from every border, no reflection can occur, offsets are compile-time constants
// and the accumulators stay in registers.  DYN = true (the two-pixel border ring of v2): offsets after reflection are run-time.
// ---------------------------------------------------------------------------------------------------------------
template <bool V2, bool DYN>
__device__ __forceinline__ void dia_row(const float* __restrict__ img, int H, int W, int y, int x, double eps, double (&acc)[DIA_P]) {
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
            double Mi[6];
            sym3_inverse_f64(M, Mi);
            const double a0 = i0 - m0, a1 = i1 - m1, a2 = i2 - m2;
            const double v0 = Mi[0] * a0 + Mi[1] * a1 + Mi[2] * a2;
            const double v1 = Mi[1] * a0 + Mi[3] * a1 + Mi[4] * a2;
            const double v2 = Mi[2] * a0 + Mi[4] * a1 + Mi[5] * a2;
#pragma unroll
            for (int py = 0; py < 3; ++py)
#pragma unroll
                for (int px = 0; px < 3; ++px) {
                    const int k = py * 3 + px;
                    const double val = -(1.0 / 9.0) - (v0 * c[k][0] + v1 * c[k][1] + v2 * c[k][2]);
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

template <bool V2>
__global__ void __launch_bounds__(128)
lap_dia_build_kernel(const float* __restrict__ img, float* __restrict__ coef, int H, int W, double eps) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y;
    if (x >= W || y >= H) return;
    double acc[DIA_P];
#pragma unroll
    for (int p = 0; p < DIA_P; ++p) acc[p] = 0.0;
    const bool border = V2 && (y < 2 || x < 2 || y >= H - 2 || x >= W - 2);
    if (border) {
        double accd[DIA_P];                 // indexed at run time (local memory); kept apart so that `acc` stays in registers
#pragma unroll
        for (int p = 0; p < DIA_P; ++p) accd[p] = 0.0;
        dia_row<V2, true>(img, H, W, y, x, eps, accd);
#pragma unroll
        for (int p = 0; p < DIA_P; ++p) acc[p] = accd[p];
    } else {
        dia_row<V2, false>(img, H, W, y, x, eps, acc);
    }
    const size_t HW = size_t(H) * W, o = size_t(y) * W + x;
#pragma unroll
    for (int p = 0; p < DIA_P; ++p) {
        const int ty = y + dia_dy(p), tx = x + dia_dx(p);
        const bool inside = ty < H && tx >= 0 && tx < W;
        coef[p * HW + o] = inside ? float(acc[p]) : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// mat-vec.  CTA = 32 x 16 output pixels, 256 threads, two rows per thread.  d = x - I of the tile plus a 2-pixel halo is
// staged in shared memory as float4 (coalesced 4-byte loads of the interleaved RGB rows); every pixel then reads its own
// 12 forward coefficients and the 12 of its backward neighbours (coalesced along x; the second use of every coefficient
// hits L1/L2), one LDS.128 per neighbour, 6 float32 operations per neighbour and channel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int DT_W = 32, DT_H = 16, DT_THREADS = 256, DT_PW = DT_W + 4, DT_PH = DT_H + 4;

__global__ void __launch_bounds__(DT_THREADS)
lap_dia_kernel(const float* __restrict__ x, const float* __restrict__ img, const float* __restrict__ LI,
               const float* __restrict__ coef, float* __restrict__ y, double* __restrict__ partial, int H, int W, float y_scale,
               int qlo, int qhi, unsigned int* __restrict__ ticket, double* __restrict__ xLx_out, const double* __restrict__ qI) {
    __shared__ float4 sd[DT_PH][DT_PW];
    __shared__ double sRed[32];
    __shared__ bool sLast;
    const int x0 = blockIdx.x * DT_W, y0 = blockIdx.y * DT_H;
    const int row_elems = W * 3;
    // ---- stage d = x - I (zero outside the image: the coefficients that point there are zero as well)
    for (int idx = threadIdx.x; idx < DT_PH * DT_PW * 3; idx += DT_THREADS) {
        const int r = idx / (DT_PW * 3), e = idx - r * (DT_PW * 3);
        const int gy = y0 - 2 + r, ge = (x0 - 2) * 3 + e;
        float v = 0.f;
        if (gy >= 0 && gy < H && ge >= 0 && ge < row_elems) {
            const size_t g = size_t(gy) * row_elems + ge;
            v = __ldg(x + g) - __ldg(img + g);
        }
        const int px = e / 3, ch = e - px * 3;
        reinterpret_cast<float*>(&sd[r][px])[ch] = v;
    }
    __syncthreads();
    const size_t HW = size_t(H) * W;
    const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;
    const int gx = x0 + tx;
    double qacc = 0.0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int ty = ty0 + half * 8, gy = y0 + ty;
        if (gx < W && gy < H) {
            const size_t o = size_t(gy) * W + gx;
            const float4 di = sd[ty + 2][tx + 2];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int p = 0; p < DIA_P; ++p) {
                const int dy = dia_dy(p), dx = dia_dx(p);
                // forward neighbour i + delta: own coefficient (stored as zero when the neighbour is outside the image)
                {
                    const float cf = __ldg(coef + p * HW + o);
                    const float4 dj = sd[ty + 2 + dy][tx + 2 + dx];
                    a0 = fmaf(cf, dj.x - di.x, a0); a1 = fmaf(cf, dj.y - di.y, a1); a2 = fmaf(cf, dj.z - di.z, a2);
                }
                // backward neighbour i - delta: its coefficient for +delta (symmetry L_ij = L_ji)
                if (gy - dy >= 0 && gx - dx >= 0 && gx - dx < W) {
                    const float cb = __ldg(coef + p * HW + o - size_t(dy) * W - dx);
                    const float4 dj = sd[ty + 2 - dy][tx + 2 - dx];
                    a0 = fmaf(cb, dj.x - di.x, a0); a1 = fmaf(cb, dj.y - di.y, a1); a2 = fmaf(cb, dj.z - di.z, a2);
                }
            }
            const float l0 = __ldg(LI + o * 3), l1 = __ldg(LI + o * 3 + 1), l2 = __ldg(LI + o * 3 + 2);
            if (partial != nullptr && gx >= qlo && gx < qhi) {
                const double x0v = double(__ldg(x + o * 3)), x1v = double(__ldg(x + o * 3 + 1)), x2v = double(__ldg(x + o * 3 + 2));
                qacc += x0v * double(a0) + double(di.x) * double(l0);
                qacc += x1v * double(a1) + double(di.y) * double(l1);
                qacc += x2v * double(a2) + double(di.z) * double(l2);
            }
            if (y != nullptr) {
                y[o * 3] = y_scale * (l0 + a0);
                y[o * 3 + 1] = y_scale * (l1 + a1);
                y[o * 3 + 2] = y_scale * (l2 + a2);
            }
        }
    }
    if (partial != nullptr) {
        // per-CTA partials summed in a fixed order by whichever CTA finishes last (deterministic, one launch, graph-replayable)
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
            a = block_sum<double>(a, sRed);
            if (threadIdx.x == 0) { *xLx_out = a + *qI; *ticket = 0u; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
bool dia_eligible(const adpst_laplacian* h) { return h->R == 1 && h->io_dtype == ADPST_F32; }

void dia_free(adpst_laplacian* h) {
    if (h->dia_coef) cudaFree(h->dia_coef);
    if (h->dia_LI) cudaFree(h->dia_LI);
    if (h->dia_qI) cudaFree(h->dia_qI);
    h->dia_coef = nullptr; h->dia_LI = nullptr; h->dia_qI = nullptr; h->dia_ready = false;
}

// I^T L I over the current quadratic window, in float64 (matrix-free kernel with x = I)
static int dia_refresh_qI(adpst_laplacian* h, cudaStream_t st) {
    const int rc = lap_matrix_free_f64(h, static_cast<const float*>(h->image), nullptr, 1.0, h->dia_qI, st);
    if (rc == ADPST_OK) h->dia_q_dirty = false;
    return rc;
}

int dia_build(adpst_laplacian* h, cudaStream_t st) {
    const size_t HW = size_t(h->H) * h->W;
    if (!h->dia_coef) ADPST_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->dia_coef), HW * DIA_P * sizeof(float)));
    if (!h->dia_LI) ADPST_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->dia_LI), HW * 3 * sizeof(float)));
    if (!h->dia_qI) ADPST_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&h->dia_qI), sizeof(double)));
    const float* img = static_cast<const float*>(h->image);
    // L I (float64 arithmetic, rounded once) and I^T L I over the whole image
    int rc = lap_matrix_free_f64(h, img, h->dia_LI, 1.0, h->dia_qI, st);
    if (rc != ADPST_OK) return rc;
    h->dia_q_dirty = (h->q_col_hi > h->q_col_lo);          // a window was set before the build: refresh on first use
    dim3 block(32, 4), grid((h->W + 31) / 32, (h->H + 3) / 4);
    if (h->mode == ADPST_LAP_V2) lap_dia_build_kernel<true><<<grid, block, 0, st>>>(img, h->dia_coef, h->H, h->W, h->eps);
    else lap_dia_build_kernel<false><<<grid, block, 0, st>>>(img, h->dia_coef, h->H, h->W, h->eps);
    ADPST_LAUNCH_CHECK();
    h->dia_ready = true;
    return ADPST_OK;
}

int dia_matvec(adpst_laplacian* h, const float* x, float* y, double y_scale, double* xLx, cudaStream_t st) {
    if (!h->dia_ready) return fail(ADPST_ERR_INVALID, "laplacian: the diagonal-format operator has not been built");
    const int qlo = h->q_col_hi > h->q_col_lo ? h->q_col_lo : 0, qhi = h->q_col_hi > h->q_col_lo ? h->q_col_hi : h->W;
    if (xLx && h->dia_q_dirty) {
        const int rc = dia_refresh_qI(h, st);
        if (rc != ADPST_OK) return rc;
    }
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
