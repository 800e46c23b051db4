// Standalone probe (compiled and run on the GPU box): LBO/SBO convention of tcgen05.mma kind::f16 for MN-major,
// 128B-swizzled operands.  Fills smem by hand in the layout gram_tc.cu uses, issues the MMAs, compares with a host
// reference.  Layout of an operand tile [MN = 128 channels][K = 32 pixels] of fp16:
//   byte offset(c, p) = (c / 64) * 4096 + (p / 8) * 1024 + (p % 8) * 128 + ((((c % 64) / 8) ^ (p % 8)) * 16) + (c % 8) * 2
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../../automated-deep-photo-style-transfer_b200/csrc/tc_common.cuh"
using namespace adpst::tc;

constexpr int M = 128, N = 128, KP = 32;

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= uint64_t((addr >> 4) & 0x3FFF);
    d |= uint64_t((lbo >> 4) & 0x3FFF) << 16;
    d |= uint64_t((sbo >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// A[c][p], B[c][p] row-major [128][KP] floats holding fp16-exact values
__global__ void probe(const float* A, const float* B, float* D, int variant) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + 8192;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < M * KP; i += blockDim.x) {
        const int c = i / KP, p = i % KP;
        const int off = (c / 64) * 4096 + (p / 8) * 1024 + (p % 8) * 128 + ((((c % 64) / 8) ^ (p % 8)) * 16) + (c % 8) * 2;
        *(__half*)(sA + off) = __float2half_rn(A[i]);
        *(__half*)(sB + off) = __float2half_rn(B[i]);
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc(&slot, 128);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_f16(M, N) | (1u << 15) | (1u << 16);
        const uint32_t lbo = variant == 0 ? 4096 : 1024, sbo = variant == 0 ? 1024 : 4096;
        for (int ks = 0; ks < KP / 16; ++ks) {
            const uint32_t a = smem_u32(sA) + ks * 2048, b = smem_u32(sB) + ks * 2048;
            umma_f16_ss(tm, desc_mn(a, lbo, sbo), desc_mn(b, lbo, sbo), idesc, ks != 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tcgen05_fence_after();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp < 4) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tm + (uint32_t(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tcgen05_fence_after(); tmem_dealloc(tm, 128); }
}

int main() {
    std::vector<float> A(M * KP), B(N * KP), D(M * N), R(M * N);
    for (auto& x : A) x = float(rand() % 17 - 8) * 0.25f;      // exactly representable in fp16
    for (auto& x : B) x = float(rand() % 13 - 6) * 0.5f;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < KP; ++k) s += double(A[m * KP + k]) * B[n * KP + k]; R[m * N + n] = float(s); }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 2048);
    for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(dD, 0, D.size() * 4);
        probe<<<1, 128, 16384 + 2048>>>(dA, dB, dD, variant);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0, mx = 0; int nz = 0;
        for (int i = 0; i < M * N; ++i) { err = fmax(err, fabs(D[i] - R[i])); mx = fmax(mx, fabs(R[i])); nz += D[i] != 0; }
        printf("variant %d: %s  max|D-R| %.4g (max|R| %.4g)  nonzero %d/%d  D[0..3]=%g %g %g %g  R[0..3]=%g %g %g %g\n", variant,
               cudaGetErrorString(e), err, mx, nz, M * N, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
