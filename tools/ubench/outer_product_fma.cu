// Microbenchmark: the accumulation loop of the structure-factor tiles (sf_fast.cu) -- per atom a thread reads 4 x-factors and
// 2 y-factors (complex, from shared memory) and adds their 4 x 2 outer product (4 real products each) to 32 real sums --
// written two ways: packed FFMA2 with the x component broadcast (the shipped form: 16 FFMA2 per atom) and scalar FFMA with
// the operands ordered for the reuse cache (32 FFMA per atom).  Question: which one does the register file / fma pipe serve
// fastest, and what is the loop's own ceiling?  (Measured: 45.7 against 47.4 cycles per atom and warp, i.e. 70 % / 67.5 % of
// the fp32 FMA peak with 4 warps per scheduler: the shipped form stays, and at C2's 39 atoms per pair the loops account for
// ~13 of the tiles kernel's ~30 us per chunk.)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/outer_product_fma tools/ubench/outer_product_fma.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pack(float x, float y) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y)); return d; }

constexpr int ATOMS = 32, TX = 64, TY = 32;

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float* out, int iters) {
    __shared__ float2 ex[ATOMS * TX];
    __shared__ float2 ey[ATOMS * TY];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    for (int i = tid; i < ATOMS * TX; i += 256) ex[i] = make_float2(1.0f + i * 1e-6f, 0.5f - i * 1e-6f);
    for (int i = tid; i < ATOMS * TY; i += 256) ey[i] = make_float2(0.25f + i * 1e-6f, 0.75f - i * 1e-6f);
    __syncthreads();
    u64 P[4][2], Q[4][2];
    float s[4][2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            P[i][kk] = Q[i][kk] = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) s[i][kk][q] = 0.f;
        }
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int a = 0; a < ATOMS; ++a) {
            if (MODE == 0) {                 // shipped: packed, x component broadcast
                u64 ys[2];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) ys[kk] = reinterpret_cast<const u64*>(ey)[a * TY + tx + 16 * kk];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 x = ex[a * TX + ty + 16 * i];
                    const u64 xc = pack(x.x, x.x), xs = pack(x.y, x.y);
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        P[i][kk] = fma2(xc, ys[kk], P[i][kk]);
                        Q[i][kk] = fma2(xs, ys[kk], Q[i][kk]);
                    }
                }
            } else {                         // scalar: 32 FFMA per atom
                float2 ys[2];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) ys[kk] = ey[a * TY + tx + 16 * kk];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 x = ex[a * TX + ty + 16 * i];
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        s[i][kk][0] = fmaf(x.x, ys[kk].x, s[i][kk][0]);
                        s[i][kk][1] = fmaf(x.x, ys[kk].y, s[i][kk][1]);
                        s[i][kk][2] = fmaf(x.y, ys[kk].x, s[i][kk][2]);
                        s[i][kk][3] = fmaf(x.y, ys[kk].y, s[i][kk][3]);
                    }
                }
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            acc += __uint_as_float((unsigned)P[i][kk]) + __uint_as_float((unsigned)(P[i][kk] >> 32)) + __uint_as_float((unsigned)Q[i][kk]) +
                   __uint_as_float((unsigned)(Q[i][kk] >> 32));
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += s[i][kk][q];
        }
    out[blockIdx.x * 256 + tid] = acc;
}

template <int MODE>
void run(const char* name, float* out, int sms, double clk_hz) {
    const int iters = 400;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<2 * sms, 256>>>(out, 10);
    cudaEventRecord(a);
    k<MODE><<<2 * sms, 256>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    // per SM: 2 CTAs x 8 warps = 16 warps, 4 per scheduler; cycles per atom and warp on one scheduler
    const double cyc = ms * 1e-3 * clk_hz / ((double)iters * ATOMS * 4);
    printf("%-52s %7.3f ms   %6.2f cycles per atom and warp (4 warps per scheduler)   %5.1f %% of the fp32 FMA peak\n", name, ms, cyc,
           100.0 * 32 / cyc);
}

int main() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * 2 * sms * 256);
    printf("SMs %d, %d MHz; 32 real FMAs per atom and thread (peak: 32 cycles per atom and warp on one scheduler)\n", sms, clk / 1000);
    run<0>("packed FFMA2, x broadcast (shipped): 16 FFMA2 + 6 LDS", out, sms, clk * 1e3);
    run<1>("scalar FFMA: 32 FFMA + 6 LDS", out, sms, clk * 1e3);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
