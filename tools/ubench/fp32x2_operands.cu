// Microbenchmark: throughput of packed fp32x2 arithmetic on sm_100a WITH DISTINCT, CHANGING REGISTER OPERANDS -- the
// situation inside an FFT butterfly -- against the loop-invariant operands of tools/ubench/fp32x2.cu (which the
// operand-reuse cache serves).  Question: is FFMA2 with three distinct 64-bit sources (six registers read per thread) a
// 2-cycle instruction like FADD2, or does the register file make it 3?  That sets the real fp32 floor of the slice step.
//   modes: 0 FFMA2 d=a*b+c (3 distinct)   1 FADD2 d=a+b (2 distinct)   2 FMUL2 d=a*b (2 distinct)
//          3 FFMA2 d=a*b+a (2 distinct)    4 FFMA2 d=a*K+c (K loop-invariant)   5 scalar FFMA 3 distinct   6 scalar FADD 2 distinct
//          7 butterfly mix: 4 FADD2 : 2 FFMA2 : 1 FMUL2 (the ratio of the row pass)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp32x2_operands tools/ubench/fp32x2_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float sfma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float sadd(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int R = 16;      // live 64-bit values per thread

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(u64* out, int iters, u64 seed) {
    u64 a[R];
    float f[2 * R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        a[i] = seed + (u64)threadIdx.x * 0x100000001ull + i;
        f[2 * i] = __uint_as_float((unsigned)(seed >> 3) + i);
        f[2 * i + 1] = __uint_as_float((unsigned)(seed >> 5) + threadIdx.x + i);
    }
    const u64 K = seed ^ 0x3f8000003f800000ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int p = (i + 5) % R, q = (i + 11) % R, r = (i + 3) % R;
            if (MODE == 0) a[i] = fma2(a[p], a[q], a[r]);
            if (MODE == 1) a[i] = add2(a[p], a[q]);
            if (MODE == 2) a[i] = mul2(a[p], a[q]);
            if (MODE == 3) a[i] = fma2(a[p], a[q], a[p]);
            if (MODE == 4) a[i] = fma2(a[p], K, a[r]);
            if (MODE == 5) { f[2 * i] = sfma(f[2 * p], f[2 * q + 1], f[2 * r]); f[2 * i + 1] = sfma(f[2 * p + 1], f[2 * q], f[2 * r + 1]); }
            if (MODE == 6) { f[2 * i] = sadd(f[2 * p], f[2 * q + 1]); f[2 * i + 1] = sadd(f[2 * p + 1], f[2 * q]); }
            if (MODE == 7) {
                const int m = i % 7;
                if (m < 4) a[i] = add2(a[p], a[q]);
                else if (m < 6) a[i] = fma2(a[p], a[q], a[r]);
                else a[i] = mul2(a[p], a[q]);
            }
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) s ^= a[i] ^ (u64)__float_as_uint(f[2 * i]) ^ ((u64)__float_as_uint(f[2 * i + 1]) << 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, u64* out, int sms, double clk_hz) {
    const int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, 512>>>(out, iters, 12345);
    cudaEventRecord(e0);
    k<MODE><<<sms, 512>>>(out, iters, 12345);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double per_thread = (double)iters * R * ((MODE == 5 || MODE == 6) ? 2 : 1);          // instructions per thread
    const double warp_instr_per_smsp = per_thread * 4;                                          // 16 warps per SM = 4 per scheduler
    const double cyc = ms * 1e-3 * clk_hz;
    printf("%-44s %8.3f ms   %.2f cycles per warp-instruction per scheduler (at %.0f MHz)\n", name, ms, cyc / warp_instr_per_smsp, clk_hz / 1e6);
}

int main() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    u64* out; cudaMalloc(&out, (size_t)sms * 512 * sizeof(u64));
    const double hz = clk * 1e3;
    run<0>("FFMA2 d=a*b+c, three distinct sources", out, sms, hz);
    run<1>("FADD2 d=a+b, two distinct sources", out, sms, hz);
    run<2>("FMUL2 d=a*b, two distinct sources", out, sms, hz);
    run<3>("FFMA2 d=a*b+a, two distinct sources", out, sms, hz);
    run<4>("FFMA2 d=a*K+c, K loop-invariant", out, sms, hz);
    run<5>("FFMA scalar, three distinct sources", out, sms, hz);
    run<6>("FADD scalar, two distinct sources", out, sms, hz);
    run<7>("mix 4 FADD2 : 2 FFMA2 : 1 FMUL2", out, sms, hz);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
