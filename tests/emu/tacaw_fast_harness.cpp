// TEST INFRASTRUCTURE (never loaded by the package): runs the stage sequence of one tile of the tiled mixed-radix TACAW
// transform (pyslice_b200/csrc/tacaw_stages.cuh, the code tacaw_fast_kernel runs between its barriers) on the host, one
// loop iteration per CUDA thread and one loop per barrier interval, so the CPU suite can check stage index arithmetic,
// twiddle selection, the digit-reversal / fftshift permutation and the tile-size rule against numpy without a GPU.
// build: g++ -std=c++20 -O2 -fPIC -DPSB_EMU -shared -o libtacaw_emu.so tacaw_fast_harness.cpp
#include "../../pyslice_b200/csrc/tacaw_stages.cuh"

#include <cmath>
#include <vector>

using namespace psb;
using namespace psb::tw;

namespace {

template <int PX, int NT>
int tile(int T, long long npix, const float2* wf, long long stride_frame, float* out) {
    constexpr int BIG = 2;
    int fac[kMaxFactors], nfac = 0;
    if (!factorise(T, fac, &nfac)) return -2;
    std::vector<float2> twd(T);
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < T; ++n) {
        const double a = -two_pi * (double)n / (double)T;
        twd[n] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    std::vector<int> perm(T);
    build_perm(T, fac, nfac, perm.data());
    const long long tiles = (npix + PX - 1) / PX;
    std::vector<float2> data((size_t)T * PX);
    for (long long b = 0; b < tiles; ++b) {
        auto each = [&](auto&& fn) {                      // one barrier interval: every thread of the CTA once
            for (unsigned tid = 0; tid < (unsigned)NT; ++tid) {
                const long long gpx = b * PX + tid % PX;
                fn(tid, gpx < npix, wf + gpx, out + gpx);
            }
        };
        if (nfac == 1) {
            each([&](unsigned tid, bool live, const float2* src, float* dst) {
                PSB_TW_DISPATCH(fac[0], (last_stage<R, PX, NT, true>(tid, data.data(), src, stride_frame, live, perm.data(), dst, npix, T)));
            });
            continue;
        }
        each([&](unsigned tid, bool live, const float2* src, float*) {
            PSB_TW_DISPATCH(fac[0], (first_stage<R, PX, NT>(tid, data.data(), src, stride_frame, live, twd.data(), T)));
        });
        int B = T / fac[0];
        for (int s = 1; s < nfac - 1; ++s) {
            each([&](unsigned tid, bool, const float2*, float*) {
                PSB_TW_DISPATCH(fac[s], (mid_stage<R, PX, NT>(tid, data.data(), twd.data(), T, B)));
            });
            B /= fac[s];
        }
        each([&](unsigned tid, bool live, const float2* src, float* dst) {
            PSB_TW_DISPATCH(fac[nfac - 1], (last_stage<R, PX, NT, false>(tid, data.data(), src, stride_frame, live, perm.data(), dst, npix, T)));
        });
    }
    return 0;
}

}  // namespace

// wf (T, npix) complex64 with frame stride `stride_frame` elements -> out (T, npix) float32; returns the tile width used
extern "C" int tacaw_fast_host(int T, long long npix, const float* wf, long long stride_frame, float* out) {
    const float2* w = reinterpret_cast<const float2*>(wf);
    const int px = pick_px(T);
    int rc = -1;
    switch (px) {
        case 64: rc = tile<64, 256>(T, npix, w, stride_frame, out); break;
        case 32: rc = tile<32, 256>(T, npix, w, stride_frame, out); break;
        case 16: rc = tile<16, 256>(T, npix, w, stride_frame, out); break;
        case 8: rc = whole_sm(T) ? tile<8, 1024>(T, npix, w, stride_frame, out) : tile<8, 256>(T, npix, w, stride_frame, out); break;
        case 4: rc = tile<4, 1024>(T, npix, w, stride_frame, out); break;
        default: return -1;
    }
    return rc == 0 ? px : rc;
}

extern "C" int tacaw_fast_plan(int T, int* fac, int* perm) {
    int nfac = 0;
    if (!factorise(T, fac, &nfac)) return -1;
    build_perm(T, fac, nfac, perm);
    return nfac;
}
