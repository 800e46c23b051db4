// runtime.cu -- error reporting and device queries shared by every entry point.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace adpst {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int current_device() {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : 0;
}

int num_sms() {
    static std::atomic<int> cached[64];            // per device: a process may drive several (different) GPUs
    const int dev = current_device() & 63;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

int device_alloc(void** p, size_t bytes, cudaStream_t st) {
    static std::atomic<unsigned long long> tuned{0};
    const int dev = current_device();
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(tuned.load(std::memory_order_acquire) & bit)) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;                 // never trim at synchronisation points
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        tuned.fetch_or(bit, std::memory_order_release);
    }
    *p = nullptr;
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(p, bytes ? bytes : 1);                // (stream capture in progress, exhausted pool, ...)
    }
    if (e != cudaSuccess) return fail(ADPST_ERR_CUDA, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return ADPST_OK;
}

void device_free(void* p, cudaStream_t st) {
    if (!p) return;
    if (cudaFreeAsync(p, st) != cudaSuccess) {               // e.g. the creating stream no longer exists
        cudaGetLastError();
        cudaFree(p);
    }
}

}  // namespace adpst

extern "C" {

int adpst_version(void) { return 100; }

const char* adpst_last_error(void) { return adpst::g_last_error.c_str(); }

unsigned long long adpst_launch_count(void) { return adpst::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
