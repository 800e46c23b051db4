// laplacian.cuh -- the matting-Laplacian handle shared by laplacian.cu (matrix-free kernels, C-ABI) and
// laplacian_dia.cu (precomputed 5x5 stencil coefficients).
#pragma once
#include "common.cuh"

struct adpst_laplacian {
    int mode, H, W, R, io_dtype, compute_dtype;
    double eps;
    void* image = nullptr;       // (H,W,3) io_dtype, owned
    double* partials = nullptr;  // one per CTA, owned; followed by the "last CTA" ticket counter
    int npartials = 0;
    int kernel = ADPST_LAP_KERNEL_AUTO;   // which mat-vec runs (adpst_laplacian_set_kernel)
    int q_col_lo = 0, q_col_hi = 0;   // x^T L x restricted to these columns (spatially tiled runs); (0,0) = all
    // diagonal-format operator (laplacian_dia.cu): r = 1, float32 storage
    float* dia_coef = nullptr;   // [12][H][W] planes: L[i, i+delta] for the 12 "forward" offsets of the 5x5 neighbourhood
    float* dia_LI = nullptr;     // [3][H][W] planes: L I, evaluated once in float64 and rounded
    double* dia_qI = nullptr;    // [W]: per-column sums of I . (L I), float64 (the constant part of x^T L x for any column window)
    bool dia_ready = false;
    cudaStream_t stream = nullptr;   // the stream the handle was created on: its buffers are released in that stream's order
};

namespace adpst {

// matrix-free float64 mat-vec (r = 1, float32 storage): y = y_scale * L x, *xLx = x^T L x over the quadratic window
int lap_matrix_free_f64(adpst_laplacian* h, const float* x, float* y, double y_scale, double* xLx, cudaStream_t st);

bool dia_eligible(const adpst_laplacian* h);
int dia_build(adpst_laplacian* h, cudaStream_t st);
void dia_free(adpst_laplacian* h);
int dia_matvec(adpst_laplacian* h, const float* x, float* y, double y_scale, double* xLx, cudaStream_t st);

}  // namespace adpst
