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
#include "tacaw_stages.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace psb {

namespace {

using namespace tw;

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

// kThreads: 256 when several tiles fit one SM's shared memory, 1024 when a tile (long series) has the SM to itself
template <int PX, int kThreads, int BIG>
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
        PSB_TW_DISPATCH(p.fac[0], (last_stage<R, PX, kThreads, true>(threadIdx.x, data, src, p.stride_frame, live, p.perm, dst, p.npix, T)));
        return;
    }
    PSB_TW_DISPATCH(p.fac[0], (first_stage<R, PX, kThreads>(threadIdx.x, data, src, p.stride_frame, live, p.tw, T)));
    __syncthreads();
    int B = T / p.fac[0];
    for (int s = 1; s < p.nfac - 1; ++s) {
        PSB_TW_DISPATCH(p.fac[s], (mid_stage<R, PX, kThreads>(threadIdx.x, data, p.tw, T, B)));
        B /= p.fac[s];
        __syncthreads();
    }
    PSB_TW_DISPATCH(p.fac[p.nfac - 1], (last_stage<R, PX, kThreads, false>(threadIdx.x, data, src, p.stride_frame, live, p.perm, dst, p.npix, T)));
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
    std::vector<int> perm(T);
    build_perm(T, tb.fac, tb.nfac, perm.data());
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

inline int radix_class(const int* fac, int nfac) {      // 0: radices up to 5, 1: up to 10, 2: up to 20
    int c = 0;
    for (int i = 0; i < nfac; ++i) c = fac[i] > 10 ? 2 : (fac[i] > 5 && c < 1 ? 1 : c);
    return c;
}

template <int PX, int kThreads, int BIG>
int go_as(const TwFastParams& p, int n_probes, cudaStream_t s) {
    const size_t smem = (size_t)p.T * PX * sizeof(float2);
    static std::atomic<size_t> smem_set[64];          // per device ordinal (the attribute is per device), zero-initialised
    const int d = rt::device() & 63;
    if (smem > smem_set[d].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(tacaw_fast_kernel<PX, kThreads, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("tacaw fast path: ") + cudaGetErrorString(e));
        smem_set[d].store(smem, std::memory_order_release);
    }
    const long long tiles = (p.npix + PX - 1) / PX;
    tacaw_fast_kernel<PX, kThreads, BIG><<<dim3((unsigned)tiles, (unsigned)n_probes), kThreads, smem, s>>>(p);
    ++launch_counter();
    return rt::check("tacaw fast launch");
}

template <int PX, int kThreads>
int go(const TwFastParams& p, int n_probes, cudaStream_t s) {
    // radix-10 / 20 butterflies want up to 127 registers: a whole-SM tile runs them with 512 threads instead of 1024
    constexpr int kBigThreads = kThreads > 512 ? 512 : kThreads;
    switch (radix_class(p.fac, p.nfac)) {
        case 2: return go_as<PX, kBigThreads, 2>(p, n_probes, s);
        case 1: return go_as<PX, kBigThreads, 1>(p, n_probes, s);
        default: return go_as<PX, kThreads, 0>(p, n_probes, s);
    }
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
    const bool one_per_sm = whole_sm(n_frames);      // 32 warps on a tile that has the SM to itself
    switch (pick_px(n_frames)) {
        case 64: {
            // radix-10 stages of T = 100 k: T/10 butterflies x 64 pixels split evenly over 320 threads but not over 256
            const long long items = (long long)(n_frames / 10) * 64;
            if (tb.fac[0] == 10 && items % 320 == 0 && items % 256 != 0) return go_as<64, 320, 1>(p, n_probes, s);
            return go<64, 256>(p, n_probes, s);
        }
        case 32: return go<32, 256>(p, n_probes, s);
        case 16: return go<16, 256>(p, n_probes, s);
        case 8: return one_per_sm ? go<8, 1024>(p, n_probes, s) : go<8, 256>(p, n_probes, s);
        case 4: return go<4, 1024>(p, n_probes, s);
        default: return fail(PSB_ERR_UNSUPPORTED, "tacaw fast path: frame count too large for a shared-memory tile");
    }
}

}  // namespace psb
