// Microbenchmark: what does HBM give for the TACAW access pattern alone?  A CTA reads a tile of PX adjacent pixels x T frames
// of complex64 (T segments of PX*8 bytes, one frame = 2 MB apart) and writes PX x T float32 (T segments of PX*4 bytes, 1 MB
// apart) -- the traffic of tacaw_fast_kernel without its transform.  Compared with a flat copy of the same 12 bytes per
// element.  If the pattern's ceiling is far below the copy bandwidth, the kernel's 0.61-0.66 of the copy peak is the layout's
// cost, not the kernel's.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tacaw_pattern tools/ubench/tacaw_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

// grid: (tiles, probes); each thread: pixel tid % PX, frames tid / PX, + step ...
template <int PX>
__global__ void __launch_bounds__(256) k_pattern(const float2* __restrict__ wf, float* __restrict__ out, int T, long long npix) {
    const int px = threadIdx.x % PX;
    constexpr int step = 256 / PX;
    const long long gpx = (long long)blockIdx.x * PX + px;
    const float2* src = wf + (long long)blockIdx.y * T * npix + gpx;
    float* dst = out + (long long)blockIdx.y * T * npix + gpx;
    for (int t = threadIdx.x / PX; t < T; t += 4 * step) {
        float2 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = t + i * step < T ? src[(long long)(t + i * step) * npix] : make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (t + i * step < T) dst[(long long)(t + i * step) * npix] = v[i].x * v[i].x + v[i].y * v[i].y;
    }
}

__global__ void __launch_bounds__(256) k_flat(const float2* __restrict__ wf, float* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float2 v = wf[i];
        out[i] = v.x * v.x + v.y * v.y;
    }
}

template <class F>
float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / 3;
}

int main() {
    const int P = 64, T = 100;
    const long long npix = 512 * 512;
    const long long n = (long long)P * T * npix;
    float2* wf;
    float* out;
    cudaMalloc(&wf, n * sizeof(float2));
    cudaMalloc(&out, n * sizeof(float));
    cudaMemset(wf, 0, n * sizeof(float2));
    const double gb = 12.0 * n / 1e9;
    printf("C3 quarter: P=%d T=%d npix=%lld, %.1f GB per pass\n", P, T, npix, gb);
    float ms = timeit([&] { k_flat<<<148 * 8, 256>>>(wf, out, n); });
    printf("flat sweep                         : %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { k_pattern<16><<<dim3((unsigned)(npix / 16), P), 256>>>(wf, out, T, npix); });
    printf("tiles of  16 pixels (128 B / 64 B) : %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { k_pattern<32><<<dim3((unsigned)(npix / 32), P), 256>>>(wf, out, T, npix); });
    printf("tiles of  32 pixels (256 B / 128 B): %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { k_pattern<64><<<dim3((unsigned)(npix / 64), P), 256>>>(wf, out, T, npix); });
    printf("tiles of  64 pixels (512 B / 256 B): %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { k_pattern<128><<<dim3((unsigned)(npix / 128), P), 256>>>(wf, out, T, npix); });
    printf("tiles of 128 pixels (1 KB / 512 B) : %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { k_pattern<256><<<dim3((unsigned)(npix / 256), P), 256>>>(wf, out, T, npix); });
    printf("tiles of 256 pixels (2 KB / 1 KB)  : %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
