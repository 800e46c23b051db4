// common.cuh -- shared helpers for libadpst (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
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

int num_sms();          // SM count of the CURRENT device (cached per device)

// Device memory owned by short-lived handles (one Laplacian per content image): stream-ordered allocations from the device's
// default memory pool, whose release threshold is raised once per device so that freed blocks stay cached.  Creating and
// destroying same-sized handles then costs no cudaMalloc / cudaFree and no device-wide synchronisation (measured: 150 ms of
// a 512x512 pair's set-up).  device_free orders the release after the work already queued on `st`.
int device_alloc(void** p, size_t bytes, cudaStream_t st);
void device_free(void* p, cudaStream_t st);
void count_launch();

// One-time per-DEVICE kernel set-up: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the current device only,
// and one process may drive several GPUs (the Python API takes device=).  `mask` is one static per kernel instantiation,
// bit d = "done on device d"; the set-up itself is idempotent, so two racing threads are harmless.
int current_device();
#define ADPST_ONCE_PER_DEVICE(stmt)                                                           \
    do {                                                                                      \
        static std::atomic<unsigned long long> _done{0};                                      \
        const unsigned long long _bit = 1ull << (::adpst::current_device() & 63);             \
        if (!(_done.load(std::memory_order_acquire) & _bit)) {                                \
            stmt;                                                                             \
            _done.fetch_or(_bit, std::memory_order_release);                                  \
        }                                                                                     \
    } while (0)

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

// Index of np.pad(..., mode='symmetric') / tf.pad(mode='SYMMETRIC'):  [.. b a | a b c | c b ..], any distance.
__device__ __forceinline__ int reflect_symmetric(int p, int n) {
    if (p >= 0 && p < n) return p;
    const int period = 2 * n;
    int q = p % period;
    if (q < 0) q += period;
    return q < n ? q : period - 1 - q;
}

}  // namespace adpst
