// Standalone probe: cycles per MMA for different orders of the 12 MMAs of one 3xTF32 stage (TS mode, M=128, N=128).
#include <cstdio>
#include "../../automated-deep-photo-style-transfer_b200/csrc/tc_common.cuh"
using namespace adpst::tc;

// BG: background activity by warps 4..7 while warp 0 issues: 0 none, 1 tcgen05.ld of the big accumulator, 2 tcgen05.st into
// the A slots, 3 LDS.128 sweeps over 16 KB, 4 STS.128 sweeps over 48 KB
template <int PAT, int BG>
__global__ void pat(long long* out, int iters, const __grid_constant__ CUtensorMap tm2d) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, bar2, tbar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) {
        unsigned h = (i + blockIdx.x * 7919u) * 2654435761u; h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        ((float*)smem)[i] = (BG == 7) ? __uint_as_float((h & 0x007FE000u) | 0x3F800000u) - 1.5f : 1.0f;   // BG 7: random tf32 operands
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1 << 20); mbar_init(&tbar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
    const uint32_t tm = slot;
    const uint32_t big = tm, small = tm + 256, a_hi = tm + 384, a_lo = tm + 416;
    if (BG == 7 && threadIdx.x < 128) {      // random A operand in tensor memory too
        uint32_t v[32]; for (int j = 0; j < 32; ++j) { unsigned h = (threadIdx.x * 33u + j + blockIdx.x) * 2654435761u; h ^= h >> 13; v[j] = (h & 0x007FE000u) | 0x3F000000u; }
        const uint32_t lb = uint32_t((threadIdx.x >> 5) * 32) << 16;
        tmem_st_32x32(tm + 384 + lb, v); tmem_st_32x32(tm + 416 + lb, v); tmem_st_32x32(tm + 448 + lb, v); tmem_st_32x32(tm + 480 + lb, v); tmem_st_wait();
        tcgen05_fence_before();
    }
    __syncthreads(); tcgen05_fence_after();
    __shared__ volatile int stop;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    long long t0 = clock64();
    if (BG == 5 && threadIdx.x == 128) {      // background TMA: 3 x 16 KB boxes per round into scratch smem
        uint32_t ph = 0;
        while (!stop) {
            mbar_arrive_expect_tx(&tbar, 3 * 16384);
            for (int j = 0; j < 3; ++j) tma_load_2d(smem + 49152 + j * 16384, &tm2d, &tbar, 0, (blockIdx.x * 3 + j) * 128);
            mbar_wait(&tbar, ph); ph ^= 1;
        }
    } else if (BG != 5 && threadIdx.x >= 128) {
        const int q = (threadIdx.x >> 5) & 3; const uint32_t lb = uint32_t(q * 32) << 16; float sink = 0.f; uint32_t v[32];
        for (int j = 0; j < 32; ++j) v[j] = j;
        while (!stop) {
            if (BG == 1) { for (int c0 = 0; c0 < 128; c0 += 32) { tmem_ld_32x32(tm + lb + c0, v); tmem_ld_wait(); sink += __uint_as_float(v[3]); } }
            if (BG == 2) { tmem_st_32x32(tm + 448 + lb, v); tmem_st_32x32(tm + 480 + lb, v); tmem_st_wait(); }
            if (BG == 3) { for (int j = 0; j < 8; ++j) { float4 x = ((float4*)smem)[j * 128 + (threadIdx.x - 128)]; sink += x.x; } }
            if (BG == 4) { for (int j = 0; j < 24; ++j) ((float4*)(smem + 49152))[j * 128 + (threadIdx.x - 128)] = make_float4(sink, 1, 2, 3); }
        }
        if (sink == 123.f) out[200] = 1;
    }
    if (threadIdx.x < 32) {
        const uint32_t idesc = umma_idesc_tf32(128, 128);
        const uint64_t dbh = umma_desc_kmajor_sw128(smem_u32(smem) + 16384, 1024), dbl = umma_desc_kmajor_sw128(smem_u32(smem) + 32768, 1024);
        for (int it = 0; it < iters; ++it) {
            if (elect_one_sync()) {
                if (PAT == 0) {            // as in conv_tc.cu: per K-step small, small, big
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_tf32_ts(small, a_lo + k * 8, dbh + k * 2, idesc, 1);
                        umma_tf32_ts(small, a_hi + k * 8, dbl + k * 2, idesc, 1);
                        umma_tf32_ts(big, a_hi + k * 8, dbh + k * 2, idesc, 1);
                    }
                } else if (PAT == 1) {     // grouped by accumulator: 4 big, then 8 small
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_tf32_ts(big, a_hi + k * 8, dbh + k * 2, idesc, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_tf32_ts(small, a_lo + k * 8, dbh + k * 2, idesc, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_tf32_ts(small, a_hi + k * 8, dbl + k * 2, idesc, 1);
                } else if (PAT == 2) {     // three different accumulators, round robin
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_tf32_ts(small, a_lo + k * 8, dbh + k * 2, idesc, 1);
                        umma_tf32_ts(small + 128, a_hi + k * 8, dbl + k * 2, idesc, 1);   // overlaps A columns on purpose? no: 384..511 -> use tm+128
                        umma_tf32_ts(big, a_hi + k * 8, dbh + k * 2, idesc, 1);
                    }
                } else {                   // all 12 into one accumulator
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_tf32_ts(big, a_lo + k * 8, dbh + k * 2, idesc, 1);
                        umma_tf32_ts(big, a_hi + k * 8, dbl + k * 2, idesc, 1);
                        umma_tf32_ts(big, a_hi + k * 8, dbh + k * 2, idesc, 1);
                    }
                }
                umma_commit(&bar2);
                umma_commit(&bar2);
            }
            __syncwarp();
        }
        if (elect_one_sync()) umma_commit(&bar);
        __syncwarp();
    }
    if (threadIdx.x < 128) mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) stop = 1;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    tcgen05_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}
template <int PAT, int BG> void run(long long* d, const CUtensorMap& tm) {
    const int iters = 400;
    cudaFuncSetAttribute(pat<PAT, BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110000);
    pat<PAT, BG><<<148, 256, 110000>>>(d, iters, tm);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
    printf("pattern %d bg %d: %s  %.1f clk per MMA (%.0f clk per 12-MMA stage)\n", PAT, BG, cudaGetErrorString(e), double(mx) / (12.0 * iters), double(mx) / iters);
}
int main() {
    long long* d; cudaMalloc(&d, 8 * 256);
    float* g; cudaMalloc(&g, size_t(148) * 3 * 128 * 32 * 4 + (1 << 20)); cudaMemset(g, 0, size_t(148) * 3 * 128 * 32 * 4);
    CUtensorMap tm; cuuint64_t dims[2] = {32, 148 * 3 * 128}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", int(r));
    run<0, 0>(d, tm); run<0, 7>(d, tm);
    return 0;
}
