// Standalone probe: how many tcgen05.mma can the issuing thread enqueue before it blocks?  (depth of the MMA issue queue)
#include <cstdio>
#include "../../automated-deep-photo-style-transfer_b200/csrc/tc_common.cuh"
using namespace adpst::tc;

template <int MODE>   // 0: SS one accumulator, 1: TS one accumulator, 2: TS small/small/big, 3: SS small/small/big
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
                for (int i = 0; i < n; ++i) {
                    const uint32_t acc = (MODE >= 2) ? ((i % 3 == 2) ? tm : tm + 256) : tm;
                    if (MODE == 1 || MODE == 2) umma_tf32_ts(acc, tm + 384 + (i & 3) * 8, db + uint64_t((i & 3) * 2), idesc, 1);
                    else umma_tf32(acc, da + uint64_t((i & 3) * 2), db + uint64_t((i & 3) * 2), idesc, 1);
                }
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
    long long h[4][64];
    cudaFuncSetAttribute(q<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000); q<0><<<1, 128, 40000>>>(d); cudaDeviceSynchronize(); cudaMemcpy(h[0], d, sizeof(h[0]), cudaMemcpyDeviceToHost);
    cudaFuncSetAttribute(q<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000); q<1><<<1, 128, 40000>>>(d); cudaDeviceSynchronize(); cudaMemcpy(h[1], d, sizeof(h[0]), cudaMemcpyDeviceToHost);
    cudaFuncSetAttribute(q<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000); q<2><<<1, 128, 40000>>>(d); cudaDeviceSynchronize(); cudaMemcpy(h[2], d, sizeof(h[0]), cudaMemcpyDeviceToHost);
    cudaFuncSetAttribute(q<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000); q<3><<<1, 128, 40000>>>(d); printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize())); cudaMemcpy(h[3], d, sizeof(h[0]), cudaMemcpyDeviceToHost);
    printf("issue time (clk) for n MMAs: SS-1acc  TS-1acc  TS-3acc  SS-3acc\n");
    for (int n = 1; n <= 16; ++n) printf("n=%2d  %6lld %6lld %6lld %6lld\n", n, h[0][n * 2], h[1][n * 2], h[2][n * 2], h[3][n * 2]);
    return 0;
}
