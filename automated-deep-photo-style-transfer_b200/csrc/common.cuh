// common.cuh -- shared helpers for libadpst (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "adpst.h"

namespace adpst {

void set_error(const std::string& msg);
int fail(int code, const char* fmt, ...);

#define ADPST_CUDA_CHECK(expr)                                                              \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::adpst::fail(ADPST_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, \
                                 cudaGetErrorString(_e));                                   \
    } while (0)

// every kernel launch of the library goes through this macro, so adpst_launch_count() is the number of OUR kernels
#define ADPST_LAUNCH_CHECK()                     \
    do {                                         \
        ::adpst::count_launch();                 \
        ADPST_CUDA_CHECK(cudaGetLastError());    \
    } while (0)

#define ADPST_REQUIRE(cond, ...)                                            \
    do {                                                                    \
        if (!(cond)) return ::adpst::fail(ADPST_ERR_INVALID, __VA_ARGS__);  \
    } while (0)

static inline cudaStream_t as_stream(adpst_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int num_sms();
void count_launch();

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in thread 0.  `scratch` must hold >= 32 elements of T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    T r = (threadIdx.x < nw) ? scratch[threadIdx.x] : T(0);
    if (wid == 0) r = warp_sum(r);
    __syncthreads();
    return r;
}

// float <-> double without F2F.  On B200 the F2F.F64.F32 / F2F.F32.F64 conversions issue on the XU pipe at about one
// lane per clock (measured with ncu: 9 conversions per row step kept the XU pipe 86 % busy and were the bottleneck of
// the float64 Laplacian stencil).  These integer versions are exact for the values that occur on this path:
//   f32_to_f64_exact: exact for zero and normal floats; float denormals (< 1.2e-38) are flushed to zero.
//   f64_to_f32_rn   : round-to-nearest-even for results in the normal float range; |d| < 2^-126 flushes to zero,
//                     |d| >= 2^128 saturates to the largest finite float (neither occurs: |y| is O(1)).
__device__ __forceinline__ double f32_to_f64_exact(float f) {
    const unsigned b = __float_as_uint(f);
    const unsigned e = (b >> 23) & 0xFFu, m = b & 0x7FFFFFu;
    unsigned hi = b & 0x80000000u, lo = 0u;
    if (e != 0u) {
        hi |= ((e + 896u) << 20) | (m >> 3);
        lo = m << 29;
    }
    return __hiloint2double(int(hi), int(lo));
}
__device__ __forceinline__ float f64_to_f32_rn(double d) {
    const unsigned hi = unsigned(__double2hiint(d)), lo = unsigned(__double2loint(d));
    const unsigned sign = hi & 0x80000000u;
    const unsigned e = (hi >> 20) & 0x7FFu;
    if (e <= 896u) return __uint_as_float(sign);                       // below the normal float range
    if (e >= 1151u) return __uint_as_float(sign | 0x7F7FFFFFu);        // above it
    const unsigned m = ((hi & 0xFFFFFu) << 3) | (lo >> 29);            // top 23 mantissa bits
    const unsigned r = lo & 0x1FFFFFFFu;                               // the 29 bits that are dropped
    const unsigned up = (r > 0x10000000u) || (r == 0x10000000u && (m & 1u));
    return __uint_as_float(sign | ((((e - 896u) << 23) + m) + up));    // a mantissa carry bumps the exponent
}

// Index of np.pad(..., mode='symmetric') / tf.pad(mode='SYMMETRIC'):  [.. b a | a b c | c b ..], any distance.
__device__ __forceinline__ int reflect_symmetric(int p, int n) {
    if (p >= 0 && p < n) return p;
    const int period = 2 * n;
    int q = p % period;
    if (q < 0) q += period;
    return q < n ? q : period - 1 - q;
}

}  // namespace adpst
