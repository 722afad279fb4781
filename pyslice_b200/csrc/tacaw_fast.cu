// TACAW time-axis transform for frame counts of the form 2^a 3^b 5^c (every BASELINE.json configuration: 20, 100, 500,
// 2000) -- reference: src/postprocessing/tacaw_data.py:61-106,
//     I[p, w, k] = | fftshift_t FFT_t( psi[p, t, k] - <psi[p, ., k]>_t ) |^2 .
//
// The generic pass (line_pass.cuh, TW) treats every pixel's time series as a strided line (element stride = one
// image) and transforms non-powers of two by Bluestein: three padded power-of-two transforms per line.  At
// configuration scale (C3: 6.7e9 elements) that is compute-bound at ~15x the 12 B/element this row needs.  Here
//
//   * a CTA owns PX adjacent pixels x all T frames: loads are T segments of PX*8 contiguous bytes, stores T segments
//     of PX*4 bytes (whole 32 B sectors for PX >= 8), the tile sits in shared memory as [t][px] in between;
//   * the transform is an in-place mixed-radix decimation-in-frequency (radices 5, 4, 3, 2 from a host factorisation):
//     stage s splits blocks of B elements into R sub-blocks, y[k1] = DFT_R(x)[k1] * W_B^(n' k1), no ping-pong buffer,
//     one CTA barrier per stage; the digit-reversed order is undone by a permutation table that also carries the
//     fftshift, so the output sweep is a plain coalesced store of |.|^2;
//   * the first stage reads its butterfly inputs straight from global memory and the last stage writes |.|^2 straight
//     back, so a T = R1 R2 R3 series crosses shared memory four times in all;
//   * psi - <psi>_t differs from psi in the zero-frequency bin only, where it is exactly 0: no mean pass, bin 0 is
//     written as 0 (the static part rides along the k = 0 chain of the decimation only; its cancellation error in the
//     other bins, <= 6e-8 * |static| * T / R_last, is of the order of the input's own float32 quantisation).
//
// Twiddles exp(-2 pi i n / T) and the permutation are built once per T in float64 on the host and cached on the device.
#include "pdl.cuh"
#include "psb_rt.h"
#include "tacaw_fast.h"

#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace psb {

namespace {

constexpr int kMaxFactors = 16;

struct TwFastParams {
    const float2* wf;
    long long stride_probe, stride_frame;   // elements
    int T;
    long long npix;
    float* out;                             // (P, T, npix)
    const float2* tw;                       // [T]  exp(-2 pi i n / T)
    const int* perm;                        // [T]  storage position -> shifted frequency index
    int nfac;
    int fac[kMaxFactors];
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }      // a * (-i)

// forward R-point DFTs in registers (W_R = exp(-2 pi i / R))
__device__ __forceinline__ void dft2(float2* x) {
    const float2 a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
}
__device__ __forceinline__ void dft3(float2* x) {
    const float c = -0.5f, s = 0.86602540378443865f;
    const float2 t1 = cadd(x[1], x[2]), t2 = csub(x[1], x[2]);
    const float2 m = make_float2(x[0].x + c * t1.x, x[0].y + c * t1.y);
    const float2 r = make_float2(s * t2.y, -s * t2.x);                 // -i*s*(x1 - x2)
    x[0] = cadd(x[0], t1);
    x[1] = cadd(m, r);
    x[2] = csub(m, r);
}
__device__ __forceinline__ void dft4(float2* x) {
    const float2 s02 = cadd(x[0], x[2]), d02 = csub(x[0], x[2]);
    const float2 s13 = cadd(x[1], x[3]), d13 = mul_mi(csub(x[1], x[3]));
    x[0] = cadd(s02, s13);
    x[2] = csub(s02, s13);
    x[1] = cadd(d02, d13);
    x[3] = csub(d02, d13);
}
__device__ __forceinline__ void dft5(float2* x) {
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;      // cos(2 pi/5), cos(4 pi/5)
    const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;       // sin(2 pi/5), sin(4 pi/5)
    const float2 a1 = cadd(x[1], x[4]), b1 = csub(x[1], x[4]);
    const float2 a2 = cadd(x[2], x[3]), b2 = csub(x[2], x[3]);
    const float2 m1 = make_float2(x[0].x + c1 * a1.x + c2 * a2.x, x[0].y + c1 * a1.y + c2 * a2.y);
    const float2 m2 = make_float2(x[0].x + c2 * a1.x + c1 * a2.x, x[0].y + c2 * a1.y + c1 * a2.y);
    const float2 r1 = make_float2(s1 * b1.y + s2 * b2.y, -(s1 * b1.x + s2 * b2.x));    // -i*(s1 b1 + s2 b2)
    const float2 r2 = make_float2(s2 * b1.y - s1 * b2.y, -(s2 * b1.x - s1 * b2.x));    // -i*(s2 b1 - s1 b2)
    x[0] = cadd(x[0], cadd(a1, a2));
    x[1] = cadd(m1, r1);
    x[4] = csub(m1, r1);
    x[2] = cadd(m2, r2);
    x[3] = csub(m2, r2);
}

template <int R>
__device__ __forceinline__ void dft(float2* x) {
    if constexpr (R == 2) dft2(x);
    else if constexpr (R == 3) dft3(x);
    else if constexpr (R == 4) dft4(x);
    else dft5(x);
}

// Stage s of the decimation in frequency splits every block of B elements into R sub-blocks of B / R:
//     y[k1 * sub + n] = DFT_R( x[. * sub + n] )[k1] * W_B^(n k1),   W_B^m = tw[m * T / B].
// The first stage (B = T) takes its inputs straight from global memory, the last one (sub = 1, no twiddles) hands
// |.|^2 straight to global memory, so an element crosses shared memory twice per middle stage and once at each end.

template <int R, int PX, int kThreads>
__device__ __forceinline__ void first_stage(float2* data, const float2* __restrict__ src, long long stride_frame, bool live,
                                            const float2* __restrict__ tw, int T) {
    const int sub = T / R;
    const int px = threadIdx.x % PX;
    for (int n = threadIdx.x / PX; n < sub; n += kThreads / PX) {
        float2 x[R];
#pragma unroll
        for (int i = 0; i < R; ++i) x[i] = live ? src[(long long)(n + i * sub) * stride_frame] : make_float2(0.f, 0.f);
        dft<R>(x);
        data[n * PX + px] = x[0];
#pragma unroll
        for (int k = 1; k < R; ++k) data[(k * sub + n) * PX + px] = cmul(x[k], __ldg(&tw[n * k]));
    }
}

template <int R, int PX, int kThreads>
__device__ __forceinline__ void mid_stage(float2* data, const float2* __restrict__ tw, int T, int B) {
    const int sub = B / R;
    const int tstep = T / B;
    const int px = threadIdx.x % PX;
    const int n_bf = T / R;
    for (int bi = threadIdx.x / PX; bi < n_bf; bi += kThreads / PX) {
        const int q = bi / sub, n = bi - q * sub;
        float2* base = data + (q * B + n) * PX + px;
        float2 x[R];
#pragma unroll
        for (int i = 0; i < R; ++i) x[i] = base[i * sub * PX];
        dft<R>(x);
        base[0] = x[0];
#pragma unroll
        for (int k = 1; k < R; ++k) base[k * sub * PX] = cmul(x[k], __ldg(&tw[n * k * tstep]));
    }
}

// kFromGlobal: the transform has a single stage (T = R), inputs come from global memory as well
template <int R, int PX, int kThreads, bool kFromGlobal>
__device__ __forceinline__ void last_stage(const float2* data, const float2* __restrict__ src, long long stride_frame, bool live,
                                           const int* __restrict__ perm, float* __restrict__ dst, long long npix, int T) {
    const int px = threadIdx.x % PX;
    const int n_bf = T / R;
    for (int q = threadIdx.x / PX; q < n_bf; q += kThreads / PX) {
        float2 x[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if (kFromGlobal) x[i] = live ? src[(long long)i * stride_frame] : make_float2(0.f, 0.f);
            else x[i] = data[(q * R + i) * PX + px];
        }
        dft<R>(x);
        if (!live) continue;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int pos = q * R + k;
            // position 0 holds X[0] = T * mean: psi - <psi>_t differs from psi in this bin only (and is 0 there)
            const float v = pos == 0 ? 0.f : x[k].x * x[k].x + x[k].y * x[k].y;
            dst[(long long)__ldg(&perm[pos]) * npix] = v;
        }
    }
}

#define PSB_TW_DISPATCH(R_, CALL)                 \
    do {                                          \
        if ((R_) == 5) { constexpr int R = 5; CALL; }      \
        else if ((R_) == 4) { constexpr int R = 4; CALL; } \
        else if ((R_) == 3) { constexpr int R = 3; CALL; } \
        else { constexpr int R = 2; CALL; }                \
    } while (0)

// kThreads: 256 when several tiles fit one SM's shared memory, 1024 when a tile (long series) has the SM to itself
template <int PX, int kThreads>
__global__ void __launch_bounds__(kThreads) tacaw_fast_kernel(const TwFastParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* data = reinterpret_cast<float2*>(smem_raw);             // [T][PX]
    const int T = p.T;
    const int px = threadIdx.x % PX;
    const long long gpx = (long long)blockIdx.x * PX + px;
    const bool live = gpx < p.npix;
    const float2* src = p.wf + (long long)blockIdx.y * p.stride_probe + gpx;
    float* dst = p.out + (long long)blockIdx.y * T * p.npix + gpx;

    if (p.nfac == 1) {
        PSB_TW_DISPATCH(p.fac[0], (last_stage<R, PX, kThreads, true>(data, src, p.stride_frame, live, p.perm, dst, p.npix, T)));
        return;
    }
    PSB_TW_DISPATCH(p.fac[0], (first_stage<R, PX, kThreads>(data, src, p.stride_frame, live, p.tw, T)));
    __syncthreads();
    int B = T / p.fac[0];
    for (int s = 1; s < p.nfac - 1; ++s) {
        PSB_TW_DISPATCH(p.fac[s], (mid_stage<R, PX, kThreads>(data, p.tw, T, B)));
        B /= p.fac[s];
        __syncthreads();
    }
    PSB_TW_DISPATCH(p.fac[p.nfac - 1], (last_stage<R, PX, kThreads, false>(data, src, p.stride_frame, live, p.perm, dst, p.npix, T)));
}

// ---- host side -------------------------------------------------------------------------------------------
struct TwTables {
    float2* tw = nullptr;
    int* perm = nullptr;
    int nfac = 0;
    int fac[kMaxFactors] = {0};
};

std::mutex g_mu;
std::map<std::pair<int, int>, TwTables> g_tables;      // (device, T)

bool factorise(int T, int* fac, int* nfac) {
    int n = 0, r = T;
    const int radices[4] = {5, 4, 3, 2};
    for (int R : radices)
        while (r % R == 0 && r > 1) {
            if (n == kMaxFactors) return false;
            fac[n++] = R;
            r /= R;
        }
    *nfac = n;
    return r == 1 && n > 0;
}

int get_tables(int T, TwTables* out, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_mu);
    const auto key = std::make_pair(rt::device(), T);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) {
        *out = it->second;
        return PSB_OK;
    }
    TwTables tb;
    if (!factorise(T, tb.fac, &tb.nfac)) return fail(PSB_ERR_UNSUPPORTED, "tacaw fast path: frame count is not 2^a 3^b 5^c");
    std::vector<float2> tw(T);
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < T; ++n) {
        const double a = -two_pi * (double)n / (double)T;
        tw[n] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    // storage position pos = k1*(T/R1) + k2*(T/(R1 R2)) + ... holds X[k], k = k1 + R1*(k2 + R2*(...)); fftshift: k -> (k + T/2) % T
    std::vector<int> perm(T);
    for (int pos = 0; pos < T; ++pos) {
        int rem = pos, B = T, k = 0, mult = 1;
        for (int i = 0; i < tb.nfac; ++i) {
            const int sub = B / tb.fac[i];
            const int d = rem / sub;
            rem -= d * sub;
            k += d * mult;
            mult *= tb.fac[i];
            B = sub;
        }
        perm[pos] = (k + T / 2) % T;
    }
    tb.tw = static_cast<float2*>(rt::dev_alloc(T * sizeof(float2)));
    tb.perm = static_cast<int*>(rt::dev_alloc(T * sizeof(int)));
    if (!tb.tw || !tb.perm) return PSB_ERR_NOMEM;
    int rc = rt::h2d(tb.tw, tw.data(), T * sizeof(float2), s);
    if (rc == PSB_OK) rc = rt::h2d(tb.perm, perm.data(), T * sizeof(int), s);
    if (rc != PSB_OK) return rc;
    g_tables[key] = tb;
    *out = tb;
    return PSB_OK;
}

template <int PX, int kThreads>
int go(const TwFastParams& p, int n_probes, cudaStream_t s) {
    const size_t smem = (size_t)p.T * PX * sizeof(float2);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(tacaw_fast_kernel<PX, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("tacaw fast path: ") + cudaGetErrorString(e));
        smem_set = smem;
    }
    const long long tiles = (p.npix + PX - 1) / PX;
    tacaw_fast_kernel<PX, kThreads><<<dim3((unsigned)tiles, (unsigned)n_probes), kThreads, smem, s>>>(p);
    ++launch_counter();
    return rt::check("tacaw fast launch");
}

// pixels per tile: as many as keep the tile within `budget` bytes of shared memory (two CTAs per SM when possible)
int pick_px(int T) {
    for (int px : {64, 32, 16, 8})
        if ((size_t)T * px * sizeof(float2) <= (96u << 10)) return px;
    if ((size_t)T * 8 * sizeof(float2) <= (200u << 10)) return 8;
    if ((size_t)T * 4 * sizeof(float2) <= (200u << 10)) return 4;
    return 0;
}

}  // namespace

bool tacaw_fast_supported(int T) {
    int fac[kMaxFactors], n = 0;
    return T >= 2 && factorise(T, fac, &n) && pick_px(T) > 0;
}

void tacaw_fast_release() {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_tables) {
        rt::dev_free(kv.second.tw);
        rt::dev_free(kv.second.perm);
    }
    g_tables.clear();
}

int launch_tacaw_fast(const float2* wf, long long stride_probe, long long stride_frame, int n_probes, int n_frames,
                      long long npix, float* intensity, cudaStream_t s) {
    TwTables tb;
    int rc = get_tables(n_frames, &tb, s);
    if (rc != PSB_OK) return rc;
    TwFastParams p;
    std::memset(&p, 0, sizeof(p));
    p.wf = wf; p.stride_probe = stride_probe; p.stride_frame = stride_frame; p.T = n_frames; p.npix = npix;
    p.out = intensity; p.tw = tb.tw; p.perm = tb.perm; p.nfac = tb.nfac;
    std::memcpy(p.fac, tb.fac, sizeof(p.fac));
    if (n_probes == 0 || npix == 0) return PSB_OK;
    const bool whole_sm = (size_t)n_frames * 8 * sizeof(float2) > (96u << 10);      // one tile per SM: 32 warps on it
    switch (pick_px(n_frames)) {
        case 64: return go<64, 256>(p, n_probes, s);
        case 32: return go<32, 256>(p, n_probes, s);
        case 16: return go<16, 256>(p, n_probes, s);
        case 8: return whole_sm ? go<8, 1024>(p, n_probes, s) : go<8, 256>(p, n_probes, s);
        case 4: return go<4, 1024>(p, n_probes, s);
        default: return fail(PSB_ERR_UNSUPPORTED, "tacaw fast path: frame count too large for a shared-memory tile");
    }
}

}  // namespace psb
