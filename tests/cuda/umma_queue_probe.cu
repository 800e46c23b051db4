// Standalone probe: how many tcgen05.mma can the issuing thread enqueue before it blocks?  (depth of the MMA issue queue)
#include <cstdio>
#include "../../automated-deep-photo-style-transfer_b200/csrc/tc_common.cuh"
using namespace adpst::tc;

__global__ void q(long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[20];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (32768) / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
    if (threadIdx.x == 0) { for (int i = 0; i < 20; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = umma_idesc_tf32(128, 128);
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem), 1024), db = umma_desc_kmajor_sw128(smem_u32(smem) + 16384, 1024);
        for (int n = 1; n <= 16; ++n) {
            long long t0 = clock64(), t1 = 0;
            if (elect_one_sync()) {
#pragma unroll 1
                for (int i = 0; i < n; ++i) umma_tf32(tm, da + uint64_t((i & 3) * 2), db + uint64_t((i & 3) * 2), idesc, 1);
                t1 = clock64();
                umma_commit(&bar[n]);
            }
            __syncwarp();
            mbar_wait(&bar[n], 0);
            long long t2 = clock64();
            if (threadIdx.x == 0 || t1) { if (t1) out[n * 2] = t1 - t0; }
            if (threadIdx.x == 0) out[n * 2 + 1] = t2 - t0;
        }
    }
    tcgen05_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}
int main() {
    long long* d; cudaMalloc(&d, 8 * 64); cudaMemset(d, 0, 8 * 64);
    cudaFuncSetAttribute(q, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    q<<<1, 128, 40000>>>(d); printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int n = 1; n <= 16; ++n) printf("n=%2d issue %5lld clk, complete %5lld clk\n", n, h[n * 2], h[n * 2 + 1]);
    return 0;
}
