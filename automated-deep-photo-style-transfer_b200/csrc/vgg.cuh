// vgg.cuh -- shared declarations of the VGG19 handle (vgg_simt.cu owns the C-ABI, conv_tc.cu the tensor-core path).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace adpst {

constexpr int kNumConv = ADPST_VGG_NUM_CONV;
// channels of the 13 convolutions block1_conv1 .. block5_conv1
__host__ __device__ constexpr int conv_cin(int i) { return i == 0 ? 3 : i <= 2 ? 64 : i <= 4 ? 128 : i <= 8 ? 256 : 512; }
__host__ __device__ constexpr int conv_cout(int i) { return i <= 1 ? 64 : i <= 3 ? 128 : i <= 7 ? 256 : 512; }

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_STYLE = 2 };
enum { CONV_PATH_TENSOR = 0, CONV_PATH_SIMT = 1 };
// max|tensor| slots of the handle (float bits, device memory): conv outputs, gradients entering conv i's data gradient,
// weights, and a scratch slot for tensors that arrive through the single-layer entry points
enum { AMAX_ACT = 0, AMAX_GRAD = 16, AMAX_WEIGHT = 32, AMAX_SCRATCH = 48, AMAX_SLOTS = 64 };

}  // namespace adpst

struct adpst_vgg {
    // exact-fp32 CUDA-core path
    float* wf[adpst::kNumConv] = {};    // [tap][Cin][Cout]
    float* wb[adpst::kNumConv] = {};    // [tap'][Cout][Cin]  (wb[0] unused)
    float* bias[adpst::kNumConv] = {};
    float* wg0 = nullptr;               // [tap'][3][64]
    // tcgen05 path: K-major FP16 hi/lo planes, index 0 = forward, 1 = data gradient; tensor maps over [9*N][K]
    void* tc_hi[2][adpst::kNumConv] = {};
    void* tc_lo[2][adpst::kNumConv] = {};
    uint32_t* amax = nullptr;           // [AMAX_SLOTS]
    CUtensorMap tm_hi[2][adpst::kNumConv];
    CUtensorMap tm_lo[2][adpst::kNumConv];
    bool tc_ready = false;
    int conv_path = adpst::CONV_PATH_TENSOR;
};

namespace adpst {
bool conv_tc_eligible(int Cin, int Cout);
int prepare_tc_weights(adpst_vgg* h, int i, cudaStream_t st);
int launch_conv_tc(adpst_vgg* h, int i, int gradient, const float* X, float* Y, const float* seed, const float* mask, int H,
                   int W, int Cin, int Cout, const uint32_t* x_absmax, uint32_t* y_absmax, float* pool_out, cudaStream_t st,
                   int pool_pitch, int pool_xoff);
// *slot = float bits of max|x| (reset: zeroes the slot first; otherwise the slot keeps the larger of its value and max|x|)
int launch_absmax(const float* x, size_t n, uint32_t* slot, cudaStream_t st, bool reset = true);
// style gradient on the tensor cores: dF[px,:] (=|+=) sum_k m_k[px]^2 F[px,:] D_k, D given as FP16 hi/lo planes (K,C,C)
// scaled by the power of two of *d_absmax; *f_absmax = max|F|
bool style_tc_eligible(int C);
int launch_style_dF_tc(const float* F, int H, int W, int C, const float* masks, int K, const void* D_hi, const void* D_lo,
                       const uint32_t* f_absmax, const uint32_t* d_absmax, float* dF, int accumulate, const void* tiles,
                       cudaStream_t st);
size_t style_tc_scratch_bytes(int HW);     // size of the per-tile class-set buffer of an H*W-pixel mask stack
int launch_style_tiles(const float* masks, int K, int H, int W, void* tiles, cudaStream_t st);
// masked Gram partials on the tensor cores (gram_tc.cu)
bool gram_tc_eligible(int C);
int gram_tc_tiles(int C);
int launch_gram_tc(const float* F, int H, int W, int C, const float* masks, int K, const int* patch_ids, const int* patch_off,
                   float* ws, int splits, const uint32_t* f_absmax, const uint32_t* m_absmax, cudaStream_t st);
}  // namespace adpst
