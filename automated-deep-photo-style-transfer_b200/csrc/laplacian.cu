// laplacian.cu -- matrix-free matting Laplacian for sm_100a.
//
// Replaces   components/matting_v2.py:11-52,147-251   (He et al. "large kernel" operator, symmetric padding)
//            components/matting_v3.py:27-102          (Levin et al. explicit COO, interior windows only)
//            components/loss.py:157-161               (x^T L x)
//
// One stencil serves both variants.  For every window k (centre pixel k, (2r+1)^2 pixels)
//     M_k = sum I I^T - s s^T / n + eps * Id        (= n * (Sigma_k + eps/n Id))
//     a_k = M_k^-1 (sum I x^T - s t^T / n),   b_k = (t - a_k^T s) / n          s = sum I, t = sum x
// and for every pixel i      y_i = cnt_i * x_i - sum_{k in w_i} (a_k^T I_i + b_k).
//   v2: windows are centred on every pixel of the symmetric-padded image; the coefficient fields are
//       symmetric-padded again (matting_v2.py:164-165).  Both pads together equal ONE symmetric pad of
//       width 2r of I and x, because a_k, b_k are invariant under reflection of the window.  cnt_i = n.
//   v3: only windows that lie fully inside the image exist (matting_v3.py:77); cnt_i = number of such windows
//       that contain i.
// Window statistics are recomputed from I every call (36 B/px of traffic instead of 72+).
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"
#include "laplacian.cuh"

namespace adpst {

// ---------------------------------------------------------------------------------------------
// per-window algebra
// ---------------------------------------------------------------------------------------------
// Inverse of the symmetric positive definite 3x3  [m0 m1 m2; m1 m3 m4; m2 m4 m5]  (same packing for the result), through its
// Cholesky factor: M = G G^T, M^-1 = G^-T G^-1.
// The cofactor / determinant formula is NOT usable here.  Where the window's colours lie on a line (grey images, two-colour
// edges) M has eigenvalues (s, eps, eps) with s/eps up to 1e6; the determinant and the cofactors then lose that many digits
// each, independently, and the result is not the inverse of any nearby matrix (measured: 1e-3 errors of L x on a grey
// image).  The Cholesky route is backward stable: the computed inverse is the exact inverse of a matrix within rounding of
// M, which is all the products M^-1 R of this file need -- like the LU behind np.linalg.inv in the reference
// (matting_v2.py:52, matting_v3.py:92).
template <typename T>
__device__ __forceinline__ void sym3_inverse(const T m[6], T inv[6]) {
    const T g00 = sqrt(m[0]);
    const T i00 = T(1) / g00;
    const T g10 = m[1] * i00, g20 = m[2] * i00;
    const T g11 = sqrt(m[3] - g10 * g10);
    const T i11 = T(1) / g11;
    const T g21 = (m[4] - g20 * g10) * i11;
    const T g22 = sqrt(m[5] - g20 * g20 - g21 * g21);
    const T i22 = T(1) / g22;
    // K = G^-1 (lower triangular): k00 = i00, k11 = i11, k22 = i22
    const T k10 = -g10 * i00 * i11;
    const T k21 = -g21 * i11 * i22;
    const T k20 = -(g20 * i00 + g21 * k10) * i22;
    // M^-1 = K^T K
    inv[0] = i00 * i00 + k10 * k10 + k20 * k20;
    inv[1] = k10 * i11 + k20 * k21;
    inv[2] = k20 * i22;
    inv[3] = i11 * i11 + k21 * k21;
    inv[4] = k21 * i22;
    inv[5] = i22 * i22;
}

// Window moments from a (2R+1)^2 patch in shared memory.  `pI`, `px` point at the patch's top-left pixel,
// `pitch` is the row pitch in elements (3 per pixel).
// Output: mu[3], pbar[3] (means), M[6] (sum of centred I I^T, + eps on the diagonal), Rc[3][3] (sum of centred
// I x^T; Rc[j][c], j = image channel, c = x channel).  float64 uses raw moments (products of float32-exact
// inputs are exact in float64); float32 centres first, which is what keeps it usable.
template <typename TIO, typename TC, int R>
__device__ __forceinline__ void window_moments(const TIO* __restrict__ pI, const TIO* __restrict__ px, int pitch,
                                               TC eps, TC mu[3], TC pbar[3], TC M[6], TC Rc[9]) {
    constexpr int D = 2 * R + 1;
    constexpr TC inv_n = TC(1) / TC(D * D);
    TC s[3] = {0, 0, 0}, t[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 6; ++i) M[i] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) Rc[i] = 0;
    if constexpr (std::is_same<TC, double>::value) {
#pragma unroll
        for (int dy = 0; dy < D; ++dy) {
#pragma unroll
            for (int dx = 0; dx < D; ++dx) {
                const TIO* qi = pI + dy * pitch + dx * 3;
                const TIO* qx = px + dy * pitch + dx * 3;
                const TC i0 = qi[0], i1 = qi[1], i2 = qi[2];
                const TC x0 = qx[0], x1 = qx[1], x2 = qx[2];
                s[0] += i0; s[1] += i1; s[2] += i2;
                t[0] += x0; t[1] += x1; t[2] += x2;
                M[0] += i0 * i0; M[1] += i0 * i1; M[2] += i0 * i2;
                M[3] += i1 * i1; M[4] += i1 * i2; M[5] += i2 * i2;
                Rc[0] += i0 * x0; Rc[1] += i0 * x1; Rc[2] += i0 * x2;
                Rc[3] += i1 * x0; Rc[4] += i1 * x1; Rc[5] += i1 * x2;
                Rc[6] += i2 * x0; Rc[7] += i2 * x1; Rc[8] += i2 * x2;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { mu[c] = s[c] * inv_n; pbar[c] = t[c] * inv_n; }
        M[0] -= s[0] * mu[0]; M[1] -= s[0] * mu[1]; M[2] -= s[0] * mu[2];
        M[3] -= s[1] * mu[1]; M[4] -= s[1] * mu[2]; M[5] -= s[2] * mu[2];
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c) Rc[j * 3 + c] -= s[j] * pbar[c];
    } else {
#pragma unroll
        for (int dy = 0; dy < D; ++dy) {
#pragma unroll
            for (int dx = 0; dx < D; ++dx) {
                const TIO* qi = pI + dy * pitch + dx * 3;
                const TIO* qx = px + dy * pitch + dx * 3;
                s[0] += TC(qi[0]); s[1] += TC(qi[1]); s[2] += TC(qi[2]);
                t[0] += TC(qx[0]); t[1] += TC(qx[1]); t[2] += TC(qx[2]);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { mu[c] = s[c] * inv_n; pbar[c] = t[c] * inv_n; }
#pragma unroll
        for (int dy = 0; dy < D; ++dy) {
#pragma unroll
            for (int dx = 0; dx < D; ++dx) {
                const TIO* qi = pI + dy * pitch + dx * 3;
                const TIO* qx = px + dy * pitch + dx * 3;
                const TC i0 = TC(qi[0]) - mu[0], i1 = TC(qi[1]) - mu[1], i2 = TC(qi[2]) - mu[2];
                const TC x0 = TC(qx[0]) - pbar[0], x1 = TC(qx[1]) - pbar[1], x2 = TC(qx[2]) - pbar[2];
                M[0] += i0 * i0; M[1] += i0 * i1; M[2] += i0 * i2;
                M[3] += i1 * i1; M[4] += i1 * i2; M[5] += i2 * i2;
                Rc[0] += i0 * x0; Rc[1] += i0 * x1; Rc[2] += i0 * x2;
                Rc[3] += i1 * x0; Rc[4] += i1 * x1; Rc[5] += i1 * x2;
                Rc[6] += i2 * x0; Rc[7] += i2 * x1; Rc[8] += i2 * x2;
            }
        }
    }
    M[0] += eps; M[3] += eps; M[5] += eps;
}

// ---------------------------------------------------------------------------------------------
// fused matvec:  y = y_scale * L x,   partial[block] = sum_tile x . (L x)
// ---------------------------------------------------------------------------------------------
template <int R> struct LapTile {
    static constexpr int TH = 16, TW = 32;                   // output pixels per CTA
    static constexpr int WH = TH + 2 * R, WW = TW + 2 * R;   // windows whose coefficients the tile needs
    static constexpr int IH = TH + 4 * R, IW = TW + 4 * R;   // input pixels those windows read
    static constexpr int NWIN = WH * WW;
    static constexpr int THREADS = 256;
    template <typename TIO, typename TC> static constexpr size_t smem_bytes() {
        return size_t(12) * NWIN * sizeof(TC) + size_t(2) * IH * IW * 3 * sizeof(TIO) + 32 * sizeof(double);
    }
};

template <typename TIO, typename TC, int R>
__global__ void __launch_bounds__(LapTile<R>::THREADS)
lap_matvec_kernel(const TIO* __restrict__ img, const TIO* __restrict__ x, TIO* __restrict__ y,
                  double* __restrict__ partial, int H, int W, int mode, TC eps, TC y_scale, int qlo, int qhi) {
    using T = LapTile<R>;
    constexpr int D = 2 * R + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TC* sC = reinterpret_cast<TC*>(smem_raw);                                  // [12][NWIN]
    double* sRed = reinterpret_cast<double*>(sC + 12 * T::NWIN);               // [32]
    TIO* sI = reinterpret_cast<TIO*>(sRed + 32);                               // [IH][IW*3]
    TIO* sX = sI + T::IH * T::IW * 3;

    const int x0 = blockIdx.x * T::TW, y0 = blockIdx.y * T::TH;
    const bool v2 = (mode == ADPST_LAP_V2);

    // phase 0: stage I and x with a 2R halo (v2: symmetric reflection; v3: zeros outside, never used)
    for (int i = threadIdx.x; i < T::IH * T::IW; i += T::THREADS) {
        const int row = i / T::IW, col = i - row * T::IW;
        int gy = y0 - 2 * R + row, gx = x0 - 2 * R + col;
        bool ok = true;
        if (v2) { gy = reflect_symmetric(gy, H); gx = reflect_symmetric(gx, W); }
        else ok = (gy >= 0 && gy < H && gx >= 0 && gx < W);
        TIO i0 = 0, i1 = 0, i2 = 0, p0 = 0, p1 = 0, p2 = 0;
        if (ok) {
            const size_t g = (size_t(gy) * W + gx) * 3;
            i0 = img[g]; i1 = img[g + 1]; i2 = img[g + 2];
            p0 = x[g];   p1 = x[g + 1];   p2 = x[g + 2];
        }
        sI[i * 3] = i0; sI[i * 3 + 1] = i1; sI[i * 3 + 2] = i2;
        sX[i * 3] = p0; sX[i * 3 + 1] = p1; sX[i * 3 + 2] = p2;
    }
    __syncthreads();

    // phase 1: coefficients (a_k: 9, b_k: 3) of every window the tile touches
    for (int w = threadIdx.x; w < T::NWIN; w += T::THREADS) {
        const int wy = w / T::WW, wx = w - wy * T::WW;
        const int cy = y0 - R + wy, cx = x0 - R + wx;                          // window centre, image coords
        const bool valid = v2 || (cy >= R && cy < H - R && cx >= R && cx < W - R);
        TC a[9], b[3];
        if (valid) {
            TC mu[3], pbar[3], M[6], Rc[9], Mi[6];
            const int off = (wy * T::IW + wx) * 3;
            window_moments<TIO, TC, R>(sI + off, sX + off, T::IW * 3, eps, mu, pbar, M, Rc);
            sym3_inverse(M, Mi);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                a[0 * 3 + c] = Mi[0] * Rc[c] + Mi[1] * Rc[3 + c] + Mi[2] * Rc[6 + c];
                a[1 * 3 + c] = Mi[1] * Rc[c] + Mi[3] * Rc[3 + c] + Mi[4] * Rc[6 + c];
                a[2 * 3 + c] = Mi[2] * Rc[c] + Mi[4] * Rc[3 + c] + Mi[5] * Rc[6 + c];
                b[c] = pbar[c] - (a[c] * mu[0] + a[3 + c] * mu[1] + a[6 + c] * mu[2]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 9; ++i) a[i] = 0;
            b[0] = b[1] = b[2] = 0;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) sC[i * T::NWIN + w] = a[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) sC[(9 + c) * T::NWIN + w] = b[c];
    }
    __syncthreads();

    // phase 2: gather the (2R+1)^2 windows that contain each output pixel
    double acc = 0.0;
    for (int o = threadIdx.x; o < T::TH * T::TW; o += T::THREADS) {
        const int oy = o / T::TW, ox = o - oy * T::TW;
        const int gy = y0 + oy, gx = x0 + ox;
        if (gy >= H || gx >= W) continue;
        TC A[9], B[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) A[i] = 0;
        B[0] = B[1] = B[2] = 0;
#pragma unroll
        for (int dy = 0; dy < D; ++dy) {
#pragma unroll
            for (int dx = 0; dx < D; ++dx) {
                const int w = (oy + dy) * T::WW + ox + dx;
#pragma unroll
                for (int i = 0; i < 9; ++i) A[i] += sC[i * T::NWIN + w];
#pragma unroll
                for (int c = 0; c < 3; ++c) B[c] += sC[(9 + c) * T::NWIN + w];
            }
        }
        TC cnt = TC(D * D);
        if (!v2) {
            const int ylo = max(gy - R, R), yhi = min(gy + R, H - R - 1);
            const int xlo = max(gx - R, R), xhi = min(gx + R, W - R - 1);
            cnt = TC(max(yhi - ylo + 1, 0) * max(xhi - xlo + 1, 0));
        }
        const int pi = ((oy + 2 * R) * T::IW + ox + 2 * R) * 3;
        const TC i0 = sI[pi], i1 = sI[pi + 1], i2 = sI[pi + 2];
        const size_t g = (size_t(gy) * W + gx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const TC xc = sX[pi + c];
            const TC yc = cnt * xc - (A[c] * i0 + A[3 + c] * i1 + A[6 + c] * i2 + B[c]);
            if (gx >= qlo && gx < qhi) acc += double(xc) * double(yc);
            if (y != nullptr) y[g + c] = TIO(y_scale * yc);
        }
    }
    if (partial != nullptr) {
        const double tot = block_sum<double>(acc, sRed);
        if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
}

// ---------------------------------------------------------------------------------------------
// r = 1 fast path: "marching warp".  The tile kernel above spends its time in shared memory (108 LDS.64 per pixel to
// gather the 9 windows of a pixel).  Here nothing goes through shared memory:
//   * a warp owns a strip of 28 output columns; lane l holds image column c0 - 2 + l;
//   * it marches down the rows keeping, in registers, a ring of the last three raw rows (own column and both
//     neighbours, refreshed with 12 shuffles per row) and a ring of the last three rows of horizontally summed window
//     coefficients (a_k: 9, b_k: 3);
//   * per row step: load one row of I and x (36 B/px of HBM traffic in total, prefetched one row ahead), compute the
//     window centred one row up from the 3x3 register patch, 3-sum its coefficients across lanes with shuffles, and
//     emit the output row two rows up as cnt*x - (A^T I + B).
// Lanes 0,1,30,31 and the first/last two rows of a strip are halo: 28/32 * RW/(RW+2) of the arithmetic is useful.
// ---------------------------------------------------------------------------------------------
constexpr int LM_COLS = 28;           // output columns per warp
constexpr int LM_WARPS = 4;           // warps per CTA

template <typename TC> __device__ __forceinline__ TC shfl_up1(TC v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <typename TC> __device__ __forceinline__ TC shfl_dn1(TC v) { return __shfl_down_sync(0xffffffffu, v, 1); }

template <typename TC>
struct LapMarchState {
    float rI[3][9], rX[3][9];        // [row slot][column (left, centre, right) * 3 + channel]
    TC hc[3][12];                     // [row slot][a (9), b (3)] summed over the three window columns
};

// One row step.  S = ring slot of the incoming raw row (compile time, so the rings stay in registers).
template <typename TC, int S>
__device__ __forceinline__ void lap_march_step(LapMarchState<TC>& st, const float (&nI)[3], const float (&nX)[3], int ir, int r0,
                                               int r_end, int gx, int lane, int H, int W, bool v2, TC eps, TC y_scale,
                                               float* __restrict__ y, double& acc, int qlo, int qhi) {
    constexpr int S1 = (S + 1) % 3, S2 = (S + 2) % 3;     // rows ir-2, ir-1 ; S holds row ir
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        st.rI[S][3 + c] = nI[c];
        st.rX[S][3 + c] = nX[c];
        st.rI[S][c] = shfl_up1(nI[c]);      st.rX[S][c] = shfl_up1(nX[c]);
        st.rI[S][6 + c] = shfl_dn1(nI[c]);  st.rX[S][6 + c] = shfl_dn1(nX[c]);
    }
    // (computed on every step: no early exits, so that no control flow the compiler cannot prove warp-uniform surrounds
    //  the shuffles; the first two steps of a strip produce unused values)
    // ---- window centred on (wr = ir - 1, gx): rows S1, S2, S of the ring
    const int wr = ir - 1;
    TC a[9], b[3];
    {
        const bool valid = (lane >= 1 && lane <= 30) && (v2 || (wr >= 1 && wr < H - 1 && gx >= 1 && gx < W - 1));
        TC mu[3], pbar[3], M[6], Rc[9];
        constexpr TC inv_n = TC(1) / TC(9);
        TC s[3] = {0, 0, 0}, t[3] = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < 6; ++i) M[i] = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) Rc[i] = 0;
        if constexpr (std::is_same<TC, double>::value) {
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const float* pi = rr == 0 ? st.rI[S1] : rr == 1 ? st.rI[S2] : st.rI[S];
                const float* px = rr == 0 ? st.rX[S1] : rr == 1 ? st.rX[S2] : st.rX[S];
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const TC i0 = pi[cc * 3], i1 = pi[cc * 3 + 1], i2 = pi[cc * 3 + 2];
                    const TC x0 = px[cc * 3], x1 = px[cc * 3 + 1], x2 = px[cc * 3 + 2];
                    s[0] += i0; s[1] += i1; s[2] += i2;
                    t[0] += x0; t[1] += x1; t[2] += x2;
                    M[0] += i0 * i0; M[1] += i0 * i1; M[2] += i0 * i2;
                    M[3] += i1 * i1; M[4] += i1 * i2; M[5] += i2 * i2;
                    Rc[0] += i0 * x0; Rc[1] += i0 * x1; Rc[2] += i0 * x2;
                    Rc[3] += i1 * x0; Rc[4] += i1 * x1; Rc[5] += i1 * x2;
                    Rc[6] += i2 * x0; Rc[7] += i2 * x1; Rc[8] += i2 * x2;
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) { mu[c] = s[c] * inv_n; pbar[c] = t[c] * inv_n; }
            M[0] -= s[0] * mu[0]; M[1] -= s[0] * mu[1]; M[2] -= s[0] * mu[2];
            M[3] -= s[1] * mu[1]; M[4] -= s[1] * mu[2]; M[5] -= s[2] * mu[2];
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) Rc[j * 3 + c] -= s[j] * pbar[c];
        } else {
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const float* pi = rr == 0 ? st.rI[S1] : rr == 1 ? st.rI[S2] : st.rI[S];
                const float* px = rr == 0 ? st.rX[S1] : rr == 1 ? st.rX[S2] : st.rX[S];
#pragma unroll
                for (int k = 0; k < 9; ++k) { s[k % 3] += pi[k]; t[k % 3] += px[k]; }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) { mu[c] = s[c] * inv_n; pbar[c] = t[c] * inv_n; }
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const float* pi = rr == 0 ? st.rI[S1] : rr == 1 ? st.rI[S2] : st.rI[S];
                const float* px = rr == 0 ? st.rX[S1] : rr == 1 ? st.rX[S2] : st.rX[S];
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const TC i0 = pi[cc * 3] - mu[0], i1 = pi[cc * 3 + 1] - mu[1], i2 = pi[cc * 3 + 2] - mu[2];
                    const TC x0 = px[cc * 3] - pbar[0], x1 = px[cc * 3 + 1] - pbar[1], x2 = px[cc * 3 + 2] - pbar[2];
                    M[0] += i0 * i0; M[1] += i0 * i1; M[2] += i0 * i2;
                    M[3] += i1 * i1; M[4] += i1 * i2; M[5] += i2 * i2;
                    Rc[0] += i0 * x0; Rc[1] += i0 * x1; Rc[2] += i0 * x2;
                    Rc[3] += i1 * x0; Rc[4] += i1 * x1; Rc[5] += i1 * x2;
                    Rc[6] += i2 * x0; Rc[7] += i2 * x1; Rc[8] += i2 * x2;
                }
            }
        }
        M[0] += eps; M[3] += eps; M[5] += eps;
        TC Mi[6];
        sym3_inverse(M, Mi);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a[c] = Mi[0] * Rc[c] + Mi[1] * Rc[3 + c] + Mi[2] * Rc[6 + c];
            a[3 + c] = Mi[1] * Rc[c] + Mi[3] * Rc[3 + c] + Mi[4] * Rc[6 + c];
            a[6 + c] = Mi[2] * Rc[c] + Mi[4] * Rc[3 + c] + Mi[5] * Rc[6 + c];
            b[c] = pbar[c] - (a[c] * mu[0] + a[3 + c] * mu[1] + a[6 + c] * mu[2]);
        }
        if (!valid) {
#pragma unroll
            for (int i = 0; i < 9; ++i) a[i] = 0;
            b[0] = b[1] = b[2] = 0;
        }
    }
    // ---- sum the coefficients of the three window columns (lanes l-1, l, l+1) into ring slot S1 (oldest, now free)
#pragma unroll
    for (int i = 0; i < 9; ++i) st.hc[S1][i] = a[i] + shfl_up1(a[i]) + shfl_dn1(a[i]);
#pragma unroll
    for (int c = 0; c < 3; ++c) st.hc[S1][9 + c] = b[c] + shfl_up1(b[c]) + shfl_dn1(b[c]);
    // ---- output row orow = ir - 2: window rows orow-1, orow, orow+1 = the three ring slots
    const int orow = ir - 2;
    if (ir >= r0 + 2 && lane >= 2 && lane <= 29 && gx < W && orow < H && orow < r_end) {
        TC cnt = TC(9);
        if (!v2) {
            const int ylo = max(orow - 1, 1), yhi = min(orow + 1, H - 2);
            const int xlo = max(gx - 1, 1), xhi = min(gx + 1, W - 2);
            cnt = TC(max(yhi - ylo + 1, 0) * max(xhi - xlo + 1, 0));
        }
        const TC i0 = st.rI[S1][3], i1 = st.rI[S1][4], i2 = st.rI[S1][5];     // raw row ir-2, own column
        const size_t g = (size_t(orow) * W + gx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const TC A0 = st.hc[0][c] + st.hc[1][c] + st.hc[2][c];
            const TC A1 = st.hc[0][3 + c] + st.hc[1][3 + c] + st.hc[2][3 + c];
            const TC A2 = st.hc[0][6 + c] + st.hc[1][6 + c] + st.hc[2][6 + c];
            const TC B = st.hc[0][9 + c] + st.hc[1][9 + c] + st.hc[2][9 + c];
            const TC xc = st.rX[S1][3 + c];
            const TC yc = cnt * xc - (A0 * i0 + A1 * i1 + A2 * i2 + B);
            if (gx >= qlo && gx < qhi) acc += double(xc) * double(yc);
            if (y != nullptr) y[g + c] = float(y_scale * yc);
        }
    }
}

__device__ __forceinline__ double shfl_up1d(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn1d(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

template <typename TC>
__global__ void __launch_bounds__(LM_WARPS * 32)
lap_march_kernel(const float* __restrict__ img, const float* __restrict__ x, float* __restrict__ y, double* __restrict__ partial,
                 int H, int W, int mode, TC eps, TC y_scale, int RW, int strips_x, int total_warps, int qlo, int qhi) {
    __shared__ double sRed[32];
    const int lane = threadIdx.x & 31;
    int gw = blockIdx.x * LM_WARPS + (threadIdx.x >> 5);
    const bool live = gw < total_warps;                        // spare warps of the last CTA run an empty strip
    if (!live) gw = 0;
    double acc = 0.0;
    {
        const int sy = gw / strips_x, sx = gw - sy * strips_x;
        const int c0 = sx * LM_COLS, r0 = sy * RW, r_end = live ? min(r0 + RW, H) : r0;
        const int gx = c0 - 2 + lane;
        const bool v2 = (mode == ADPST_LAP_V2);
        const int mx = v2 ? reflect_symmetric(gx, W) : gx;
        const bool col_ok = v2 || (gx >= 0 && gx < W);
        LapMarchState<TC> st;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 12; ++j) st.hc[i][j] = 0;
#pragma unroll
            for (int j = 0; j < 9; ++j) { st.rI[i][j] = 0.f; st.rX[i][j] = 0.f; }
        }
        auto load_row = [&](int ir, float (&vI)[3], float (&vX)[3]) {
            int my = ir;
            bool ok = col_ok && live;
            if (v2) my = reflect_symmetric(ir, H);
            else ok = ok && ir >= 0 && ir < H;
            vI[0] = vI[1] = vI[2] = vX[0] = vX[1] = vX[2] = 0.f;
            if (ok) {
                const size_t g = (size_t(my) * W + mx) * 3;
                vI[0] = __ldg(img + g); vI[1] = __ldg(img + g + 1); vI[2] = __ldg(img + g + 2);
                vX[0] = __ldg(x + g);   vX[1] = __ldg(x + g + 1);   vX[2] = __ldg(x + g + 2);
            }
        };
        float cI[3], cX[3], nI[3], nX[3];
        const int ir_begin = r0 - 2;
        const int ntriples = (RW + 4 + 2) / 3;                      // same trip count for every warp (see lap_march2_kernel)
        load_row(ir_begin, cI, cX);
        for (int tpl = 0; tpl < ntriples; ++tpl) {
            const int ir = ir_begin + 3 * tpl;
            load_row(ir + 1, nI, nX);                               // prefetch one row ahead
            lap_march_step<TC, 0>(st, cI, cX, ir, r0, r_end, gx, lane, H, W, v2, eps, y_scale, y, acc, qlo, qhi);
            load_row(ir + 2, cI, cX);
            lap_march_step<TC, 1>(st, nI, nX, ir + 1, r0, r_end, gx, lane, H, W, v2, eps, y_scale, y, acc, qlo, qhi);
            load_row(ir + 3, nI, nX);
            lap_march_step<TC, 2>(st, cI, cX, ir + 2, r0, r_end, gx, lane, H, W, v2, eps, y_scale, y, acc, qlo, qhi);
#pragma unroll
            for (int c = 0; c < 3; ++c) { cI[c] = nI[c]; cX[c] = nX[c]; }
        }
    }
    if (partial != nullptr) {
        const double tot = block_sum<double>(acc, sRed);
        if (threadIdx.x == 0) partial[blockIdx.x] = tot;
    }
}

// ---------------------------------------------------------------------------------------------
// float64 marching warp ("march3"): what runs for r = 1, float32 I/O, float64 arithmetic.
// The kernel is bound by the float64 pipe (64 lanes/clk/SM), not by HBM, so the design minimises float64 instructions
// and shuffles per pixel:
//   * every lane owns TWO adjacent columns: a 3-wide horizontal sum of a field costs 3 adds and 2 shuffles per column
//     pair (q = v0 + v1, W0 = q + left neighbour's v1, W1 = q + right neighbour's v0) instead of 4 adds and 4 shuffles,
//     the strip is 64 columns wide (60 useful: 94 %), and the two windows give the scheduler independent work;
//   * the 12 coefficient fields are not kept for three window rows; a finished window row is contracted with the image
//     at once and ADDED into the three output rows it touches (z[o] += Ha^T I_o + Hb), so the carried state per column
//     is 9 doubles instead of 36;
//   * 1/n, eps and the validity mask ride on operations that exist anyway (eps/3 is the addend of the first diagonal
//     product of each column, b is carried as n*b and divided in the final FMA, invalid windows -- v3 only -- are selected
//     away with integer moves);
//   * float <-> double conversions use the hardware F2F (18 per column pair and row; the XU pipe they issue on is at 1 %):
//     an integer emulation costs ~100 issue slots per pixel and was measured 12 % slower.
// Lanes 0 and 31 and the first/last two rows of a strip are halo.
// ---------------------------------------------------------------------------------------------
constexpr int L3_COLS = 60;           // output columns per warp
constexpr int L3_WARPS = 4;           // warps per CTA

struct Lap3State {
    double rI[3][2][3], rX[3][2][3];  // [row slot][column of the pair][channel]
    double z[3][2][3];                // [output row slot][column][x channel]: sum over windows of a^T I + b, so far
};

__device__ __forceinline__ double sel_f64(bool keep, double v) {   // keep ? v : 0, without the float64 pipe
    return __hiloint2double(keep ? __double2hiint(v) : 0, keep ? __double2loint(v) : 0);
}

template <int S, bool V2>
__device__ __forceinline__ void lap_march3_step(Lap3State& st, const float (&nI)[6], const float (&nX)[6], int ir, int r0,
                                                int r_end, int gx0, int lane, int H, int W, double eps3,
                                                double y_scale, float* __restrict__ y, double& acc, int qlo, int qhi) {
    constexpr bool v2 = V2;
    constexpr int S1 = (S + 1) % 3, S2 = (S + 2) % 3;     // slots of rows ir-2, ir-1 (S holds row ir)
    constexpr double inv_n = 1.0 / 9.0;
#pragma unroll
    for (int col = 0; col < 2; ++col)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            st.rI[S][col][c] = double(nI[col * 3 + c]);
            st.rX[S][col][c] = double(nX[col * 3 + c]);
        }
    // ---- raw moments of each own column over rows ir-2..ir:  s(3) t(3) Q(6) R(9); exact in float64
    double cm[2][21];
#pragma unroll
    for (int col = 0; col < 2; ++col) {
        {
            const double i0 = st.rI[S1][col][0], i1 = st.rI[S1][col][1], i2 = st.rI[S1][col][2];
            const double x0 = st.rX[S1][col][0], x1 = st.rX[S1][col][1], x2 = st.rX[S1][col][2];
            double* m = cm[col];
            m[0] = i0; m[1] = i1; m[2] = i2; m[3] = x0; m[4] = x1; m[5] = x2;
            m[6] = fma(i0, i0, eps3); m[7] = i0 * i1; m[8] = i0 * i2; m[9] = fma(i1, i1, eps3); m[10] = i1 * i2;
            m[11] = fma(i2, i2, eps3);
            m[12] = i0 * x0; m[13] = i0 * x1; m[14] = i0 * x2;
            m[15] = i1 * x0; m[16] = i1 * x1; m[17] = i1 * x2;
            m[18] = i2 * x0; m[19] = i2 * x1; m[20] = i2 * x2;
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int slot = rr == 0 ? S2 : S;
            const double i0 = st.rI[slot][col][0], i1 = st.rI[slot][col][1], i2 = st.rI[slot][col][2];
            const double x0 = st.rX[slot][col][0], x1 = st.rX[slot][col][1], x2 = st.rX[slot][col][2];
            double* m = cm[col];
            m[0] += i0; m[1] += i1; m[2] += i2; m[3] += x0; m[4] += x1; m[5] += x2;
            m[6] = fma(i0, i0, m[6]); m[7] = fma(i0, i1, m[7]); m[8] = fma(i0, i2, m[8]);
            m[9] = fma(i1, i1, m[9]); m[10] = fma(i1, i2, m[10]); m[11] = fma(i2, i2, m[11]);
            m[12] = fma(i0, x0, m[12]); m[13] = fma(i0, x1, m[13]); m[14] = fma(i0, x2, m[14]);
            m[15] = fma(i1, x0, m[15]); m[16] = fma(i1, x1, m[16]); m[17] = fma(i1, x2, m[17]);
            m[18] = fma(i2, x0, m[18]); m[19] = fma(i2, x1, m[19]); m[20] = fma(i2, x2, m[20]);
        }
    }
    // ---- windows centred on (wr = ir-1, gx0) and (wr, gx0+1): columns gx0-1..gx0+1 and gx0..gx0+2
    const int wr = ir - 1;
    double cf[2][12];                                          // a (9: [image channel j][x channel c] at j*3+c), n*b (3)
    {
        double wm[2][21];
#pragma unroll
        for (int i = 0; i < 21; ++i) {
            const double q = cm[0][i] + cm[1][i];
            wm[0][i] = q + shfl_up1d(cm[1][i]);
            wm[1][i] = q + shfl_dn1d(cm[0][i]);
        }
#pragma unroll
        for (int col = 0; col < 2; ++col) {
            const double* w = wm[col];
            const int gx = gx0 + col;
            const bool valid = v2 || (wr >= 1 && wr < H - 1 && gx >= 1 && gx < W - 1);
            const double mu0 = w[0] * inv_n, mu1 = w[1] * inv_n, mu2 = w[2] * inv_n;
            double M[6], Rc[9], Mi[6];
            M[0] = fma(-w[0], mu0, w[6]); M[1] = fma(-w[0], mu1, w[7]); M[2] = fma(-w[0], mu2, w[8]);
            M[3] = fma(-w[1], mu1, w[9]); M[4] = fma(-w[1], mu2, w[10]); M[5] = fma(-w[2], mu2, w[11]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                Rc[c] = fma(-mu0, w[3 + c], w[12 + c]);
                Rc[3 + c] = fma(-mu1, w[3 + c], w[15 + c]);
                Rc[6 + c] = fma(-mu2, w[3 + c], w[18 + c]);
            }
            sym3_inverse(M, Mi);
            double* a = cf[col];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double a0 = fma(Mi[2], Rc[6 + c], fma(Mi[1], Rc[3 + c], Mi[0] * Rc[c]));
                const double a1 = fma(Mi[4], Rc[6 + c], fma(Mi[3], Rc[3 + c], Mi[1] * Rc[c]));
                const double a2 = fma(Mi[5], Rc[6 + c], fma(Mi[4], Rc[3 + c], Mi[2] * Rc[c]));
                const double nb = fma(-a2, w[2], fma(-a1, w[1], fma(-a0, w[0], w[3 + c])));      // n*b = t - a^T s
                if (V2) {
                    a[c] = a0; a[3 + c] = a1; a[6 + c] = a2; a[9 + c] = nb;          // every window exists (reflection)
                } else {
                    a[c] = sel_f64(valid, a0); a[3 + c] = sel_f64(valid, a1); a[6 + c] = sel_f64(valid, a2);
                    a[9 + c] = sel_f64(valid, nb);
                }
            }
        }
    }
    // ---- coefficients summed over the three window columns around each own column
    double hc[2][12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const double q = cf[0][i] + cf[1][i];
        hc[0][i] = q + shfl_up1d(cf[1][i]);
        hc[1][i] = q + shfl_dn1d(cf[0][i]);
    }
    // ---- window row wr touches output rows wr-1 (slot S1, complete after this), wr (slot S2), wr+1 (slot S, first term)
#pragma unroll
    for (int col = 0; col < 2; ++col) {
        const double* h = hc[col];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double hb = h[9 + c] * inv_n;
            st.z[S1][col][c] = fma(h[6 + c], st.rI[S1][col][2], fma(h[3 + c], st.rI[S1][col][1], fma(h[c], st.rI[S1][col][0], st.z[S1][col][c] + hb)));
            st.z[S2][col][c] = fma(h[6 + c], st.rI[S2][col][2], fma(h[3 + c], st.rI[S2][col][1], fma(h[c], st.rI[S2][col][0], st.z[S2][col][c] + hb)));
            st.z[S][col][c] = fma(h[6 + c], st.rI[S][col][2], fma(h[3 + c], st.rI[S][col][1], fma(h[c], st.rI[S][col][0], hb)));
        }
    }
    // ---- output row orow = ir - 2
    const int orow = ir - 2;
    if (ir >= r0 + 2 && lane >= 1 && lane <= 30 && orow < H && orow < r_end) {
#pragma unroll
        for (int col = 0; col < 2; ++col) {
            const int gx = gx0 + col;
            if (gx < W) {
                double cnt = 9.0;
                if (!v2) {
                    const int ylo = max(orow - 1, 1), yhi = min(orow + 1, H - 2);
                    const int xlo = max(gx - 1, 1), xhi = min(gx + 1, W - 2);
                    const int nwin = max(yhi - ylo + 1, 0) * max(xhi - xlo + 1, 0);      // 0..9, table instead of I2F
                    cnt = nwin == 9 ? 9.0 : nwin == 6 ? 6.0 : nwin == 4 ? 4.0 : nwin == 3 ? 3.0 : nwin == 2 ? 2.0 : nwin == 1 ? 1.0 : 0.0;
                }
                const size_t g = (size_t(orow) * W + gx) * 3;
                const bool inq = gx >= qlo && gx < qhi;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double xc = st.rX[S1][col][c];
                    const double yc = fma(cnt, xc, -st.z[S1][col][c]);
                    if (inq) acc = fma(xc, yc, acc);
                    if (y != nullptr) y[g + c] = __double2float_rn(y_scale * yc);
                }
            }
        }
    }
}

template <bool V2>
__global__ void __launch_bounds__(L3_WARPS * 32)
lap_march3_kernel(const float* __restrict__ img, const float* __restrict__ x, float* __restrict__ y, double* __restrict__ partial,
                  int H, int W, double eps, double y_scale, int RW, int strips_x, int total_warps, int qlo, int qhi,
                  unsigned int* __restrict__ ticket, double* __restrict__ xLx_out) {
    __shared__ double sRed[32];
    __shared__ bool sLast;
    const int lane = threadIdx.x & 31;
    int gw = blockIdx.x * L3_WARPS + (threadIdx.x >> 5);
    const bool live = gw < total_warps;                        // spare warps of the last CTA run an empty strip
    if (!live) gw = 0;
    double acc = 0.0;
    {
        const int sy = gw / strips_x, sx = gw - sy * strips_x;
        const int c0 = sx * L3_COLS, r0 = sy * RW, r_end = live ? min(r0 + RW, H) : r0;
        const int gx0 = c0 - 2 + 2 * lane;
        constexpr bool v2 = V2;
        int mx[2];
        bool col_ok[2];
#pragma unroll
        for (int col = 0; col < 2; ++col) {
            const int gx = gx0 + col;
            mx[col] = v2 ? reflect_symmetric(gx, W) : gx;
            col_ok[col] = live && (v2 || (gx >= 0 && gx < W));
        }
        Lap3State st;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int col = 0; col < 2; ++col)
#pragma unroll
                for (int j = 0; j < 3; ++j) { st.rI[i][col][j] = 0.0; st.rX[i][col][j] = 0.0; st.z[i][col][j] = 0.0; }
        auto load_row = [&](int ir, float (&vI)[6], float (&vX)[6]) {
            int my = ir;
            bool row_ok = true;
            if (v2) my = reflect_symmetric(ir, H);
            else row_ok = ir >= 0 && ir < H;
#pragma unroll
            for (int col = 0; col < 2; ++col) {
                vI[col * 3] = vI[col * 3 + 1] = vI[col * 3 + 2] = vX[col * 3] = vX[col * 3 + 1] = vX[col * 3 + 2] = 0.f;
                if (row_ok && col_ok[col]) {
                    const size_t g = (size_t(my) * W + mx[col]) * 3;
                    vI[col * 3] = __ldg(img + g); vI[col * 3 + 1] = __ldg(img + g + 1); vI[col * 3 + 2] = __ldg(img + g + 2);
                    vX[col * 3] = __ldg(x + g);   vX[col * 3 + 1] = __ldg(x + g + 1);   vX[col * 3 + 2] = __ldg(x + g + 2);
                }
            }
        };
        float cI[6], cX[6], nI[6], nX[6];
        const int ir_begin = r0 - 2;
        // rows r0-2 .. r0+RW+1 for every warp: the trip count depends on kernel parameters only, so the shuffles stay
        // outside divergence handling; rows past the image / the strip are predicated off at the store
        const int ntriples = (RW + 4 + 2) / 3;
        const double eps3 = eps * (1.0 / 3.0);
        load_row(ir_begin, cI, cX);
        for (int tpl = 0; tpl < ntriples; ++tpl) {
            const int ir = ir_begin + 3 * tpl;
            load_row(ir + 1, nI, nX);                               // one row ahead
            lap_march3_step<0, V2>(st, cI, cX, ir, r0, r_end, gx0, lane, H, W, eps3, y_scale, y, acc, qlo, qhi);
            load_row(ir + 2, cI, cX);
            lap_march3_step<1, V2>(st, nI, nX, ir + 1, r0, r_end, gx0, lane, H, W, eps3, y_scale, y, acc, qlo, qhi);
            load_row(ir + 3, nI, nX);
            lap_march3_step<2, V2>(st, cI, cX, ir + 2, r0, r_end, gx0, lane, H, W, eps3, y_scale, y, acc, qlo, qhi);
#pragma unroll
            for (int c = 0; c < 6; ++c) { cI[c] = nI[c]; cX[c] = nX[c]; }
        }
    }
    if (partial != nullptr) {
        // x^T L x: per-CTA partials, summed in a fixed order by whichever CTA finishes last (deterministic, no second launch;
        // the ticket counter is left at zero, so the kernel can be replayed from a CUDA graph)
        const double tot = block_sum<double>(acc, sRed);
        if (threadIdx.x == 0) {
            partial[blockIdx.x] = tot;
            __threadfence();
            sLast = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        }
        __syncthreads();
        if (sLast) {
            __threadfence();
            double a = 0.0;
            for (int i = threadIdx.x; i < int(gridDim.x); i += blockDim.x) a += partial[i];
            a = block_sum<double>(a, sRed);
            if (threadIdx.x == 0) { *xLx_out = a; *ticket = 0u; }
        }
    }
}

__global__ void sum_partials_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += partial[i];   // fixed order: deterministic
    a = block_sum<double>(a, red);
    if (threadIdx.x == 0) *out = a;
}

// ---------------------------------------------------------------------------------------------
// v2 coefficient fields (matting_v2.py:49-52): means = s/n, delta_inv = (Sigma + eps/n Id)^-1 = n M^-1
// ---------------------------------------------------------------------------------------------
template <typename TIO, typename TC, int R>
__global__ void lap_coeff_kernel(const TIO* __restrict__ img, TIO* __restrict__ means, TIO* __restrict__ dinv,
                                 int H, int W, TC eps) {
    constexpr int D = 2 * R + 1;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y * blockDim.y + threadIdx.y;
    if (gx >= W || gy >= H) return;
    TC s[3] = {0, 0, 0}, Q[6] = {0, 0, 0, 0, 0, 0};
    TC v[D * D][3];
#pragma unroll
    for (int dy = 0; dy < D; ++dy) {
        const int yy = reflect_symmetric(gy - R + dy, H);
#pragma unroll
        for (int dx = 0; dx < D; ++dx) {
            const int xx = reflect_symmetric(gx - R + dx, W);
            const size_t g = (size_t(yy) * W + xx) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) { v[dy * D + dx][c] = TC(img[g + c]); s[c] += v[dy * D + dx][c]; }
        }
    }
    constexpr TC inv_n = TC(1) / TC(D * D);
    const TC mu[3] = {s[0] * inv_n, s[1] * inv_n, s[2] * inv_n};
#pragma unroll
    for (int k = 0; k < D * D; ++k) {
        const TC a = v[k][0] - mu[0], b = v[k][1] - mu[1], c = v[k][2] - mu[2];
        Q[0] += a * a; Q[1] += a * b; Q[2] += a * c; Q[3] += b * b; Q[4] += b * c; Q[5] += c * c;
    }
    Q[0] += eps; Q[3] += eps; Q[5] += eps;
    TC Mi[6];
    sym3_inverse(Q, Mi);
    const TC n = TC(D * D);
    const size_t p = size_t(gy) * W + gx;
#pragma unroll
    for (int c = 0; c < 3; ++c) means[p * 3 + c] = TIO(mu[c]);
    TIO* o = dinv + p * 9;
    o[0] = TIO(n * Mi[0]); o[1] = TIO(n * Mi[1]); o[2] = TIO(n * Mi[2]);
    o[3] = TIO(n * Mi[1]); o[4] = TIO(n * Mi[3]); o[5] = TIO(n * Mi[4]);
    o[6] = TIO(n * Mi[2]); o[7] = TIO(n * Mi[4]); o[8] = TIO(n * Mi[5]);
}

// ---------------------------------------------------------------------------------------------
// v3 COO export (matting_v3.py:87-100): window-major, then (a, b) row-major; duplicates kept.
//   vals[a][b] = delta_ab - (1/n) (1 + (I_a - mu)^T (Sigma + eps/n Id)^-1 (I_b - mu))
// ---------------------------------------------------------------------------------------------
template <typename TIO, typename TC, int R>
__global__ void __launch_bounds__(256)
lap_export_coo_kernel(const TIO* __restrict__ img, int64_t* __restrict__ rows, int64_t* __restrict__ cols,
                      TIO* __restrict__ vals, int H, int W, TC eps) {
    constexpr int D = 2 * R + 1, N = D * D, WPB = (R == 3 ? 8 : 32);  // windows per block
    __shared__ TC sD[WPB][N][3];                                  // centred pixels
    __shared__ TC sXv[WPB][N][3];                                 // (Sigma + eps/n)^-1 d_a
    const int cw = W - 2 * R, ch = H - 2 * R;
    const long long nwin = (long long)cw * ch;
    const long long w0 = (long long)blockIdx.x * WPB;
    static_assert(sizeof(TC) * WPB * N * 3 * 2 <= 48 * 1024, "static shared memory budget");
    if (threadIdx.x < WPB && w0 + threadIdx.x < nwin) {
        const long long k = w0 + threadIdx.x;
        const int wy = int(k / cw), wx = int(k - (long long)wy * cw);
        TC s[3] = {0, 0, 0};
        TC v[N][3];
#pragma unroll
        for (int a = 0; a < N; ++a) {
            const size_t g = (size_t(wy + a / D) * W + wx + a % D) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) { v[a][c] = TC(img[g + c]); s[c] += v[a][c]; }
        }
        constexpr TC inv_n = TC(1) / TC(N);
        TC Q[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int a = 0; a < N; ++a) {
#pragma unroll
            for (int c = 0; c < 3; ++c) v[a][c] -= s[c] * inv_n;
            Q[0] += v[a][0] * v[a][0]; Q[1] += v[a][0] * v[a][1]; Q[2] += v[a][0] * v[a][2];
            Q[3] += v[a][1] * v[a][1]; Q[4] += v[a][1] * v[a][2]; Q[5] += v[a][2] * v[a][2];
        }
        Q[0] += eps; Q[3] += eps; Q[5] += eps;
        TC Mi[6];
        sym3_inverse(Q, Mi);                                      // (Sigma + eps/n)^-1 = n * Mi
#pragma unroll
        for (int a = 0; a < N; ++a) {
            const TC d0 = v[a][0], d1 = v[a][1], d2 = v[a][2];
            sD[threadIdx.x][a][0] = d0; sD[threadIdx.x][a][1] = d1; sD[threadIdx.x][a][2] = d2;
            sXv[threadIdx.x][a][0] = TC(N) * (Mi[0] * d0 + Mi[1] * d1 + Mi[2] * d2);
            sXv[threadIdx.x][a][1] = TC(N) * (Mi[1] * d0 + Mi[3] * d1 + Mi[4] * d2);
            sXv[threadIdx.x][a][2] = TC(N) * (Mi[2] * d0 + Mi[4] * d1 + Mi[5] * d2);
        }
    }
    __syncthreads();
    const int nloc = int(min((long long)WPB, nwin - w0));
    constexpr TC inv_n = TC(1) / TC(N);
    for (int e = threadIdx.x; e < nloc * N * N; e += blockDim.x) {
        const int lw = e / (N * N), ab = e - lw * N * N, a = ab / N, b = ab - a * N;
        const long long k = w0 + lw;
        const int wy = int(k / cw), wx = int(k - (long long)wy * cw);
        const TC q = sXv[lw][a][0] * sD[lw][b][0] + sXv[lw][a][1] * sD[lw][b][1] + sXv[lw][a][2] * sD[lw][b][2];
        const TC val = (a == b ? TC(1) : TC(0)) - inv_n * (TC(1) + q);
        const size_t o = size_t(k) * N * N + ab;
        rows[o] = int64_t(wy + a / D) * W + wx + a % D;
        cols[o] = int64_t(wy + b / D) * W + wx + b % D;
        vals[o] = TIO(val);
    }
}

}  // namespace adpst

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
namespace adpst {

// rows per marching warp: enough warps for ~2 waves of 8 warps per SM, at most 64 rows (halo overhead (RW+2)/RW)
static inline int march_rows(int H, int W, int warps_per_sm, int cols = LM_COLS) {
    const int strips = (W + cols - 1) / cols;
    const int want = warps_per_sm * num_sms();   // taller strips waste fewer halo rows, more warps hide more latency
    int rw = 64;
    while (rw > 16 && strips * ((H + rw - 1) / rw) < want) rw /= 2;
    return rw;
}

template <typename TC>
static int launch_march(adpst_laplacian* h, const void* x, void* y, double y_scale, double* xLx, cudaStream_t st) {
    // float64 kernel: 255 registers -> 8 resident warps per SM, one full wave; float32 kernel: 12+ resident, two waves
    const int qlo = h->q_col_hi > h->q_col_lo ? h->q_col_lo : 0, qhi = h->q_col_hi > h->q_col_lo ? h->q_col_hi : h->W;
    constexpr bool f64 = std::is_same<TC, double>::value;
    const int cols = f64 ? L3_COLS : LM_COLS;
    const int RW = march_rows(h->H, h->W, f64 ? 8 : 16, cols);
    const int strips_x = (h->W + cols - 1) / cols, total = strips_x * ((h->H + RW - 1) / RW);
    const int ctas = (total + LM_WARPS - 1) / LM_WARPS;
    if (ctas > h->npartials) return fail(ADPST_ERR_INVALID, "laplacian: partial buffer too small (%d > %d)", ctas, h->npartials);
    if constexpr (f64) {
        const float* img = static_cast<const float*>(h->image);
        const float* xf = static_cast<const float*>(x);
        float* yf = static_cast<float*>(y);
        double* part = xLx ? h->partials : nullptr;
        unsigned int* ticket = reinterpret_cast<unsigned int*>(h->partials + h->npartials);
        if (h->mode == ADPST_LAP_V2)
            lap_march3_kernel<true><<<ctas, L3_WARPS * 32, 0, st>>>(img, xf, yf, part, h->H, h->W, h->eps, y_scale, RW, strips_x,
                                                                    total, qlo, qhi, ticket, xLx);
        else
            lap_march3_kernel<false><<<ctas, L3_WARPS * 32, 0, st>>>(img, xf, yf, part, h->H, h->W, h->eps, y_scale, RW, strips_x,
                                                                     total, qlo, qhi, ticket, xLx);
        ADPST_LAUNCH_CHECK();
    } else {
        lap_march_kernel<TC><<<ctas, LM_WARPS * 32, 0, st>>>(static_cast<const float*>(h->image), static_cast<const float*>(x),
                                                             static_cast<float*>(y), xLx ? h->partials : nullptr, h->H, h->W,
                                                             h->mode, TC(h->eps), TC(y_scale), RW, strips_x, total, qlo, qhi);
        ADPST_LAUNCH_CHECK();
        if (xLx) {
            sum_partials_kernel<<<1, 256, 0, st>>>(h->partials, ctas, xLx);
            ADPST_LAUNCH_CHECK();
        }
    }
    return ADPST_OK;
}

template <typename TIO, typename TC, int R>
static int launch_matvec(adpst_laplacian* h, const void* x, void* y, double y_scale, double* xLx, cudaStream_t st) {
    if constexpr (R == 1 && std::is_same<TIO, float>::value) {
        if (h->kernel != ADPST_LAP_KERNEL_TILE) return launch_march<TC>(h, x, y, y_scale, xLx, st);
    }
    using T = LapTile<R>;
    const int qlo = h->q_col_hi > h->q_col_lo ? h->q_col_lo : 0, qhi = h->q_col_hi > h->q_col_lo ? h->q_col_hi : h->W;
    auto kern = lap_matvec_kernel<TIO, TC, R>;
    const size_t smem = T::template smem_bytes<TIO, TC>();
    ADPST_ONCE_PER_DEVICE(ADPST_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))));
    dim3 grid((h->W + T::TW - 1) / T::TW, (h->H + T::TH - 1) / T::TH);
    kern<<<grid, T::THREADS, smem, st>>>(static_cast<const TIO*>(h->image), static_cast<const TIO*>(x),
                                          static_cast<TIO*>(y), xLx ? h->partials : nullptr, h->H, h->W, h->mode,
                                          TC(h->eps), TC(y_scale), qlo, qhi);
    ADPST_LAUNCH_CHECK();
    if (xLx) {
        sum_partials_kernel<<<1, 256, 0, st>>>(h->partials, int(grid.x * grid.y), xLx);
        ADPST_LAUNCH_CHECK();
    }
    return ADPST_OK;
}

template <typename TIO, typename TC, int R>
static int launch_coeffs(adpst_laplacian* h, void* means, void* dinv, cudaStream_t st) {
    dim3 block(32, 8), grid((h->W + 31) / 32, (h->H + 7) / 8);
    lap_coeff_kernel<TIO, TC, R><<<grid, block, 0, st>>>(static_cast<const TIO*>(h->image), static_cast<TIO*>(means),
                                                         static_cast<TIO*>(dinv), h->H, h->W, TC(h->eps));
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

template <typename TIO, typename TC, int R>
static int launch_export(adpst_laplacian* h, int64_t* rows, int64_t* cols, void* vals, cudaStream_t st) {
    const long long nwin = (long long)(h->W - 2 * R) * (h->H - 2 * R);
    if (nwin <= 0) return ADPST_OK;
    constexpr int WPB = (R == 3 ? 8 : 32);
    const unsigned blocks = unsigned((nwin + WPB - 1) / WPB);
    lap_export_coo_kernel<TIO, TC, R><<<blocks, 256, 0, st>>>(static_cast<const TIO*>(h->image), rows, cols,
                                                              static_cast<TIO*>(vals), h->H, h->W, TC(h->eps));
    ADPST_LAUNCH_CHECK();
    return ADPST_OK;
}

// dispatch over (io dtype, compute dtype, radius)
#define ADPST_LAP_DISPATCH(FN, h, ...)                                                              \
    do {                                                                                            \
        const int key = (h)->io_dtype * 2 + (h)->compute_dtype;                                     \
        switch ((h)->R) {                                                                           \
            case 1:                                                                                 \
                if (key == 0) return FN<float, float, 1>(h, __VA_ARGS__);                           \
                if (key == 1) return FN<float, double, 1>(h, __VA_ARGS__);                          \
                if (key == 3) return FN<double, double, 1>(h, __VA_ARGS__);                         \
                break;                                                                              \
            case 2:                                                                                 \
                if (key == 0) return FN<float, float, 2>(h, __VA_ARGS__);                           \
                if (key == 1) return FN<float, double, 2>(h, __VA_ARGS__);                          \
                if (key == 3) return FN<double, double, 2>(h, __VA_ARGS__);                         \
                break;                                                                              \
            case 3:                                                                                 \
                if (key == 0) return FN<float, float, 3>(h, __VA_ARGS__);                           \
                if (key == 1) return FN<float, double, 3>(h, __VA_ARGS__);                          \
                if (key == 3) return FN<double, double, 3>(h, __VA_ARGS__);                         \
                break;                                                                              \
        }                                                                                           \
        return fail(ADPST_ERR_UNSUPPORTED, "laplacian: unsupported (io=%d, compute=%d, radius=%d)", \
                    (h)->io_dtype, (h)->compute_dtype, (h)->R);                                     \
    } while (0)

static int dispatch_matvec(adpst_laplacian* h, const void* x, void* y, double ys, double* xLx, cudaStream_t st) {
    if (h->kernel == ADPST_LAP_KERNEL_DIA || (h->kernel == ADPST_LAP_KERNEL_AUTO && h->dia_ready))
        return dia_matvec(h, static_cast<const float*>(x), static_cast<float*>(y), ys, xLx, st);
    ADPST_LAP_DISPATCH(launch_matvec, h, x, y, ys, xLx, st);
}

int lap_matrix_free_f64(adpst_laplacian* h, const float* x, float* y, double y_scale, double* xLx, cudaStream_t st) {
    return launch_march<double>(h, x, y, y_scale, xLx, st);
}
static int dispatch_coeffs(adpst_laplacian* h, void* means, void* dinv, cudaStream_t st) {
    ADPST_LAP_DISPATCH(launch_coeffs, h, means, dinv, st);
}
static int dispatch_export(adpst_laplacian* h, int64_t* rows, int64_t* cols, void* vals, cudaStream_t st) {
    ADPST_LAP_DISPATCH(launch_export, h, rows, cols, vals, st);
}

}  // namespace adpst

extern "C" {

int adpst_laplacian_create(int mode, int H, int W, int radius, double epsilon, const void* image_dev, int io_dtype,
                           int compute_dtype, adpst_stream_t stream, adpst_laplacian** out) {
    using namespace adpst;
    ADPST_REQUIRE(out != nullptr, "laplacian_create: out is NULL");
    *out = nullptr;
    ADPST_REQUIRE(mode == ADPST_LAP_V2 || mode == ADPST_LAP_V3, "laplacian_create: mode must be V2 or V3");
    ADPST_REQUIRE(H >= 1 && W >= 1, "laplacian_create: empty image (%d x %d)", H, W);
    ADPST_REQUIRE(image_dev != nullptr, "laplacian_create: image is NULL");
    ADPST_REQUIRE(epsilon > 0.0, "laplacian_create: epsilon must be > 0");
    ADPST_REQUIRE(io_dtype == ADPST_F32 || io_dtype == ADPST_F64, "laplacian_create: bad io_dtype");
    ADPST_REQUIRE(compute_dtype == ADPST_F32 || compute_dtype == ADPST_F64, "laplacian_create: bad compute_dtype");
    if (radius < 1 || radius > 3)
        return fail(ADPST_ERR_UNSUPPORTED, "laplacian_create: window_radius %d not supported (1..3)", radius);
    if (io_dtype == ADPST_F64 && compute_dtype == ADPST_F32)
        return fail(ADPST_ERR_UNSUPPORTED, "laplacian_create: float64 I/O with float32 arithmetic is not provided");
    auto* h = new adpst_laplacian();
    h->mode = mode; h->H = H; h->W = W; h->R = radius; h->io_dtype = io_dtype; h->compute_dtype = compute_dtype;
    h->eps = epsilon;
    const size_t bytes = size_t(H) * W * 3 * (io_dtype == ADPST_F64 ? 8 : 4);
    h->stream = as_stream(stream);
    cudaError_t e = device_alloc(&h->image, bytes, h->stream) == ADPST_OK ? cudaSuccess : cudaErrorMemoryAllocation;
    if (e == cudaSuccess) {
        h->npartials = ((W + 31) / 32) * ((H + 15) / 16) + 64;
        if (device_alloc(reinterpret_cast<void**>(&h->partials), sizeof(double) * (h->npartials + 1), h->stream) != ADPST_OK)
            e = cudaErrorMemoryAllocation;
        if (e == cudaSuccess) e = cudaMemsetAsync(h->partials + h->npartials, 0, sizeof(double), as_stream(stream));
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->image, image_dev, bytes, cudaMemcpyDeviceToDevice, as_stream(stream));
    if (e != cudaSuccess) {
        adpst_laplacian_destroy(h);
        return fail(ADPST_ERR_CUDA, "laplacian_create: %s", cudaGetErrorString(e));
    }
    if (dia_eligible(h)) {               // r = 1, float32 storage: precompute the 5x5 stencil coefficients (the hot path)
        const int rc = dia_build(h, as_stream(stream));
        if (rc != ADPST_OK) {
            adpst_laplacian_destroy(h);
            return rc;
        }
    }
    *out = h;
    return ADPST_OK;
}

int adpst_laplacian_set_kernel(adpst_laplacian* h, int kind, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(h != nullptr, "laplacian_set_kernel: NULL handle");
    ADPST_REQUIRE(kind >= ADPST_LAP_KERNEL_AUTO && kind <= ADPST_LAP_KERNEL_TILE, "laplacian_set_kernel: unknown kernel %d", kind);
    if (kind == ADPST_LAP_KERNEL_DIA) {
        if (!dia_eligible(h))
            return fail(ADPST_ERR_UNSUPPORTED, "laplacian_set_kernel: the diagonal-format kernel needs radius 1 and float32 storage");
        if (!h->dia_ready) {
            const int rc = dia_build(h, as_stream(stream));
            if (rc != ADPST_OK) return rc;
        }
    }
    h->kernel = kind;
    return ADPST_OK;
}

int adpst_laplacian_kernel(const adpst_laplacian* h) {
    if (!h) return -1;
    if (h->kernel != ADPST_LAP_KERNEL_AUTO) return h->kernel;
    return h->dia_ready ? ADPST_LAP_KERNEL_DIA : ADPST_LAP_KERNEL_MATRIX_FREE;
}

void adpst_laplacian_destroy(adpst_laplacian* h) {
    if (!h) return;
    adpst::device_free(h->image, h->stream);
    adpst::device_free(h->partials, h->stream);
    adpst::dia_free(h);
    delete h;
}

int adpst_laplacian_matvec(adpst_laplacian* h, const void* x_dev, void* y_dev, double y_scale, double* xLx_dev,
                           adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(h != nullptr && x_dev != nullptr, "laplacian_matvec: NULL handle or x");
    ADPST_REQUIRE(y_dev != nullptr || xLx_dev != nullptr, "laplacian_matvec: nothing to compute (y and xLx NULL)");
    return dispatch_matvec(h, x_dev, y_dev, y_scale, xLx_dev, as_stream(stream));
}

int adpst_laplacian_coefficients(adpst_laplacian* h, void* means_dev, void* delta_inv_dev, adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(h != nullptr && means_dev != nullptr && delta_inv_dev != nullptr, "laplacian_coefficients: NULL argument");
    if (h->mode != ADPST_LAP_V2)
        return fail(ADPST_ERR_UNSUPPORTED, "laplacian_coefficients: only the v2 operator has means/delta_inv fields");
    return dispatch_coeffs(h, means_dev, delta_inv_dev, as_stream(stream));
}

int adpst_laplacian_set_quadratic_window(adpst_laplacian* h, int col_lo, int col_hi) {
    using namespace adpst;
    ADPST_REQUIRE(h != nullptr && col_lo >= 0 && col_hi <= h->W && col_lo <= col_hi, "laplacian_set_quadratic_window: bad window");
    h->q_col_lo = col_lo;
    h->q_col_hi = col_hi;
    return ADPST_OK;
}

int64_t adpst_laplacian_nnz(const adpst_laplacian* h) {
    if (!h) return -1;
    const int64_t d = 2 * h->R + 1, ch = h->H - 2 * h->R, cw = h->W - 2 * h->R;
    if (ch <= 0 || cw <= 0) return 0;
    return d * d * d * d * ch * cw;
}

int adpst_laplacian_export_coo(adpst_laplacian* h, int64_t* rows_dev, int64_t* cols_dev, void* vals_dev,
                               adpst_stream_t stream) {
    using namespace adpst;
    ADPST_REQUIRE(h != nullptr, "laplacian_export_coo: NULL handle");
    if (h->mode != ADPST_LAP_V3)
        return fail(ADPST_ERR_UNSUPPORTED, "laplacian_export_coo: only the v3 operator has an explicit matrix");
    if (adpst_laplacian_nnz(h) == 0) return ADPST_OK;
    ADPST_REQUIRE(rows_dev && cols_dev && vals_dev, "laplacian_export_coo: NULL output");
    return dispatch_export(h, rows_dev, cols_dev, vals_dev, as_stream(stream));
}

}  // extern "C"
