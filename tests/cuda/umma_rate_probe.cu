// Standalone probe: issue rate of tcgen05.mma kind::tf32 (cycles per MMA) for M=128, N in {64,128,256}, A from smem (SS) or TMEM (TS),
// one CTA per SM on all SMs, operands resident in shared memory (no TMA traffic).
#include <cstdio>
#include <vector>
#include "../../automated-deep-photo-style-transfer_b200/csrc/tc_common.cuh"
using namespace adpst::tc;

// pattern: 0 = alternate two accumulators, 1 = always the same accumulator, 2 = the 3xTF32 sequence of conv_tc.cu
// (small, small, big per K-step; commit to an mbarrier every 12 MMAs)
template <int N, bool TS>
__global__ void rate(long long* out, int n_mma, int pattern = 0) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1 << 20); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
    const uint32_t tm = slot;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        const uint32_t idesc = umma_idesc_tf32(128, N);
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem), 1024), db = umma_desc_kmajor_sw128(smem_u32(smem) + 16384, 1024);
        t0 = clock64();
        if (elect_one_sync()) {
            for (int i = 0; i < n_mma; ++i) {
                const uint64_t koff = uint64_t((i & 3) * 2);
                uint32_t acc = tm + (i & 1) * N;
                if (pattern == 1) acc = tm;
                if (pattern == 2 || pattern == 3) acc = (i % 3 == 2) ? tm + ((i / 24) & 1) * N : tm + 2 * N;
                if (pattern == 4) acc = tm;
                if (pattern == 5 || pattern == 6) acc = (i % 12 >= 8) ? tm + ((i / 24) & 1) * N : tm + 2 * N;
                if (TS) umma_tf32_ts(acc, tm + 448 + (i & 3) * 8, db + koff, idesc, 1);
                else umma_tf32(acc, da + koff, db + koff, idesc, 1);
                if ((pattern == 2 || pattern == 4 || pattern == 6) && i % 12 == 11) umma_commit(&bar2);
                if (pattern == 7 && i % 12 == 11) { umma_commit(&bar2); umma_commit(&bar2); umma_commit(&bar2); }
            }
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    tcgen05_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, bool TS> void run(const char* name, long long* d, int sms) {
    const int n = 4096, smem = 16384 + N * 128 + 2048;
    cudaFuncSetAttribute(rate<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int pattern : {0, 2, 3, 4, 5, 6, 7}) {
        const int grid = sms;
        rate<N, TS><<<grid, 128, smem>>>(d, n, pattern);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        printf("%s N=%d pattern=%d: %s  %.1f clk/MMA  (%.0f MAC/clk/SM)\n", name, N, pattern, cudaGetErrorString(e), double(mx) / n,
               128.0 * N * 8 * n / double(mx));
    }
}

int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d; cudaMalloc(&d, 8 * 1024);
    run<128, false>("SS", d, sms); run<128, true>("TS", d, sms);
    return 0;
}
