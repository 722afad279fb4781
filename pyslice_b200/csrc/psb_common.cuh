// Common definitions for the pyslice_b200 CUDA engine (sm_100a).
//
// Every kernel body in this directory is written as a functor
//     struct K { static constexpr int kThreads, kMinBlocks;
//                template <class Ctx> static PSB_D void run(const Ctx& cx, const Params& p); };
// and launched through psb::launch<K>(grid, smem, stream, params).  `Ctx` hides threadIdx /
// blockIdx / __syncthreads / dynamic shared memory.  The product build (nvcc) only ever
// instantiates DevCtx.  Defining PSB_EMU (done only by tests/emu/, compiled with g++) swaps in a
// thread-per-CUDA-thread CPU emulation so the `-m "not gpu"` tests can exercise the very same
// index arithmetic without a device; the package never loads that build.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>

#ifdef PSB_EMU
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdlib>
#include <thread>
#include <vector>

#define PSB_HD inline
#define PSB_D inline
#define PSB_RESTRICT
struct float2 { float x, y; };
struct double2 { double x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef void* cudaStream_t;
static inline void sincospif(float x, float* s, float* c) {
    const double a = 3.14159265358979323846 * (double)x;
    *s = (float)std::sin(a); *c = (float)std::cos(a);
}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __int2float_rn(int v) { return (float)v; }
#else
#include <atomic>
#include <cuda_runtime.h>
#define PSB_HD __host__ __device__ __forceinline__
#define PSB_D __device__ __forceinline__
#define PSB_RESTRICT __restrict__
#endif

namespace psb {

// number of kernel launches issued by this library (bench.py reports it as gpu_launches)
inline std::atomic<long long>& launch_counter() {
    static std::atomic<long long> n{0};
    return n;
}

// ---- complex helpers (float2 = complex64) -------------------------------------------------
PSB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
PSB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
PSB_HD float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
PSB_HD float2 cmulc(float2 a, float2 b) {
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
PSB_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
PSB_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// ---- execution contexts -------------------------------------------------------------------
#ifndef PSB_EMU
struct DevCtx {
    PSB_D int tid() const { return threadIdx.x; }
    PSB_D int bx() const { return blockIdx.x; }
    PSB_D int by() const { return blockIdx.y; }
    PSB_D int bz() const { return blockIdx.z; }
    PSB_D int gx() const { return gridDim.x; }
    PSB_D void sync() const { __syncthreads(); }
    PSB_D unsigned char* smem() const {
        extern __shared__ __align__(16) unsigned char psb_dyn_smem[];
        return psb_dyn_smem;
    }
    PSB_D int atomic_add(int* p, int v) const { return atomicAdd(p, v); }
    PSB_D float atomic_add(float* p, float v) const { return atomicAdd(p, v); }
};

template <class K, class P>
__global__ void __launch_bounds__(K::kThreads, K::kMinBlocks) psb_kernel(const P p) {
    DevCtx cx;
    K::run(cx, p);
}

// Launch kernel functor K on `stream`.  Returns cudaGetLastError().
template <class K, class P>
inline cudaError_t launch(dim3 grid, size_t smem, cudaStream_t stream, const P& p) {
    if (smem > 48 * 1024) {
        // the opt-in is a per-device function attribute: remember what was granted per device ordinal
        static std::atomic<size_t> granted[64];   // per instantiation, zero-initialised
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        if (smem > granted[d].load(std::memory_order_acquire)) {
            cudaError_t e = cudaFuncSetAttribute(psb_kernel<K, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            granted[d].store(smem, std::memory_order_release);
        }
    }
    psb_kernel<K, P><<<grid, K::kThreads, smem, stream>>>(p);
    ++launch_counter();
    return cudaGetLastError();
}
#else
typedef int cudaError_t;
static const int cudaSuccess = 0;
struct EmuCtx {
    int tid_, bx_, by_, bz_, gx_;
    std::barrier<>* bar_;
    unsigned char* smem_;
    int tid() const { return tid_; }
    int bx() const { return bx_; }
    int by() const { return by_; }
    int bz() const { return bz_; }
    int gx() const { return gx_; }
    void sync() const { bar_->arrive_and_wait(); }
    unsigned char* smem() const { return smem_; }
    int atomic_add(int* p, int v) const { return std::atomic_ref<int>(*p).fetch_add(v); }
    float atomic_add(float* p, float v) const {
        std::atomic_ref<float> r(*p);
        float old = r.load();
        while (!r.compare_exchange_weak(old, old + v)) {}
        return old;
    }
};

template <class K, class P>
inline cudaError_t launch(dim3 grid, size_t smem, cudaStream_t, const P& p) {
    const int nt = K::kThreads;
    std::vector<unsigned char> sm(smem + 64);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                std::barrier<> bar(nt);
                std::vector<std::thread> th;
                th.reserve(nt);
                for (int t = 0; t < nt; ++t)
                    th.emplace_back([&, t] {
                        EmuCtx cx{t, (int)bx, (int)by, (int)bz, (int)grid.x, &bar, sm.data()};
                        K::run(cx, p);
                    });
                for (auto& x : th) x.join();
            }
    return 0;
}
#endif

}  // namespace psb
