// Microbenchmark: the stage exchange of a 256-point line FFT held by 16 lanes x 16 registers (thread j has positions
// (j, e), e = 0..15, and needs (e, j) for the next butterfly layer: a 16 x 16 transposition of complex64 values inside a
// half warp) done (a) through shared memory the way the fused kernels do it (16 STS.64 into a padded buffer, __syncwarp,
// 16 LDS.64) and (b) with warp shuffles only (four xor stages, each swapping half of the registers with the partner lane:
// 16 SHFL.b32 + selects per stage).  The north star asks for warp-shuffle butterflies; this prices them.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/exchange tools/ubench/exchange_shfl_vs_smem.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;

__device__ __forceinline__ u64 shfl_xor64(u64 v, int mask) {
    unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return ((u64)hi << 32) | lo;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(u64* out, int iters) {
    extern __shared__ u64 sm_all[];                      // per warp: two lines of 256 + 16 padding
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = lane >> 4, j = lane & 15;
    u64 v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = ((u64)threadIdx.x << 32) | (unsigned)(e * 7 + blockIdx.x);
    u64* buf = sm_all + warp * 2 * 272 + c * 272;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int e = 0; e < 16; ++e) { const int q = 16 * j + e; buf[q + (q >> 4)] = v[e]; }
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 16; ++e) { const int q = 16 * e + j; v[e] = buf[q + (q >> 4)]; }
            __syncwarp();
        } else {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int bit = 1 << s;
                const bool up = (j & bit) != 0;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    if (e & bit) continue;
                    const u64 a = v[e], b = v[e | bit];
                    const u64 recv = shfl_xor64(up ? a : b, bit);
                    v[e] = up ? recv : a;
                    v[e | bit] = up ? b : recv;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] += (u64)e;     // keep the iterations dependent and the values live
    }
    u64 s = 0;
#pragma unroll
    for (int e = 0; e < 16; ++e) s ^= v[e];
    out[blockIdx.x * 512 + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, u64* out, int sms, double clk_hz, u64* check) {
    const int iters = 2000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int smem = 16 * 2 * 272 * 8;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<MODE><<<sms, 512, smem>>>(out, 3);
    cudaMemcpy(check, out, 8, cudaMemcpyDeviceToHost);
    cudaEventRecord(a);
    k<MODE><<<sms, 512, smem>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    // 16 warps per SM = 4 per scheduler; SM cycles per exchange of one warp (two lines of 256 points)
    printf("%-46s %7.3f ms   %6.1f SM cycles per warp exchange (16 warps resident)   %5.2f cycles per complex value\n", name, ms,
           ms * 1e-3 * clk_hz / ((double)iters * 16), ms * 1e-3 * clk_hz / ((double)iters * 16 * 512));
}

int main() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    u64* out;
    cudaMalloc(&out, sizeof(u64) * sms * 512);
    u64 c0 = 0, c1 = 0;
    printf("SMs %d, %d MHz; 16 x 16 transposition of complex64 inside each half warp, 512 values per warp\n", sms, clk / 1000);
    run<0>("shared memory (16 STS.64 + 16 LDS.64, padded)", out, sms, clk * 1e3, &c0);
    run<1>("warp shuffles (4 xor stages x 16 SHFL.b32)", out, sms, clk * 1e3, &c1);
    printf("same result: %s\nstatus: %s\n", c0 == c1 ? "yes" : "NO", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
