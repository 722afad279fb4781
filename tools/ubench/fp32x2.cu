// Microbenchmark: issue throughput of scalar FFMA/FADD vs packed FFMA2/FADD2 on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/build/fp32x2 tools/ubench/fp32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 b = make_float2(s, 1.0f - s), c = make_float2(1e-3f, 2e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
            if (MODE == 1) a[i] = fma2(a[i], b, c);
            if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }
            if (MODE == 3) a[i] = add2(a[i], c);
            if (MODE == 4) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = a[i].y + c.y; }   // FFMA + FADD mix
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, float* out) {
    const int iters = 4096, blocks = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, iters, 0.999f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 0.999f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = (double)blocks * 256 * iters * 16;   // fp32 element operations
    printf("%-22s %8.3f ms  %8.2f T elem-ops/s  (%.1f per clk per SM at 1.9 GHz)\n", name, ms, lane_ops / ms * 1e-9,
           lane_ops / (ms * 1e-3) / 148 / 1.9e9);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    run<0>("FFMA scalar", out);
    run<1>("FFMA2 packed", out);
    run<2>("FADD scalar", out);
    run<3>("FADD2 packed", out);
    run<4>("FFMA+FADD mix", out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
