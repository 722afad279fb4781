// Register butterflies, stages and host-side planning of the tiled mixed-radix TACAW transform (tacaw_fast.cu).
// Device code in the product build; the same source compiles for the host (PSB_EMU) so tests/emu/tacaw_fast_harness.cpp
// can run a tile's stage sequence thread by thread and check index arithmetic, twiddle selection and the output
// permutation against numpy without a GPU.
#pragma once
#include "psb_common.cuh"

#include <cstdlib>

namespace psb {
namespace tw {

constexpr int kMaxFactors = 16;

PSB_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }      // a * (-i)

// forward R-point DFTs in registers (W_R = exp(-2 pi i / R))
PSB_HD void dft2(float2* x) {
    const float2 a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
}
PSB_HD void dft3(float2* x) {
    const float c = -0.5f, s = 0.86602540378443865f;
    const float2 t1 = cadd(x[1], x[2]), t2 = csub(x[1], x[2]);
    const float2 m = make_float2(x[0].x + c * t1.x, x[0].y + c * t1.y);
    const float2 r = make_float2(s * t2.y, -s * t2.x);                 // -i*s*(x1 - x2)
    x[0] = cadd(x[0], t1);
    x[1] = cadd(m, r);
    x[2] = csub(m, r);
}
PSB_HD void dft4(float2* x) {
    const float2 s02 = cadd(x[0], x[2]), d02 = csub(x[0], x[2]);
    const float2 s13 = cadd(x[1], x[3]), d13 = mul_mi(csub(x[1], x[3]));
    x[0] = cadd(s02, s13);
    x[2] = csub(s02, s13);
    x[1] = cadd(d02, d13);
    x[3] = csub(d02, d13);
}
PSB_HD void dft5(float2* x) {
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
PSB_HD void dft(float2* x) {
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
PSB_D void first_stage(unsigned tid, float2* data, const float2* PSB_RESTRICT src, long long stride_frame, bool live,
                                            const float2* PSB_RESTRICT tw, int T) {
    const int sub = T / R;
    const int px = tid % PX;
    constexpr int step = kThreads / PX;
    // Two butterflies per trip: the 2*R global loads are all issued before the first butterfly is computed.  With one
    // butterfly per trip a CTA kept R*kThreads*8 bytes in flight -- about half of what the HBM latency-bandwidth product
    // asks of an SM (ncu r1j: long-scoreboard stalls 4.0 per issue, 0.61 of the copy bandwidth).
    int n = tid / PX;
    for (; n + step < sub; n += 2 * step) {
        float2 x[R], y[R];
#pragma unroll
        for (int i = 0; i < R; ++i) x[i] = live ? src[(long long)(n + i * sub) * stride_frame] : make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < R; ++i) y[i] = live ? src[(long long)(n + step + i * sub) * stride_frame] : make_float2(0.f, 0.f);
        dft<R>(x);
        data[n * PX + px] = x[0];
#pragma unroll
        for (int k = 1; k < R; ++k) data[(k * sub + n) * PX + px] = cmul(x[k], __ldg(&tw[n * k]));
        dft<R>(y);
        data[(n + step) * PX + px] = y[0];
#pragma unroll
        for (int k = 1; k < R; ++k) data[(k * sub + n + step) * PX + px] = cmul(y[k], __ldg(&tw[(n + step) * k]));
    }
    for (; n < sub; n += step) {
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
PSB_D void mid_stage(unsigned tid, float2* data, const float2* PSB_RESTRICT tw, int T, int B) {
    const int sub = B / R;
    const int tstep = T / B;
    const int px = tid % PX;
    const int n_bf = T / R;
    for (int bi = tid / PX; bi < n_bf; bi += kThreads / PX) {
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
PSB_D void last_stage(unsigned tid, const float2* data, const float2* PSB_RESTRICT src, long long stride_frame, bool live,
                                           const int* PSB_RESTRICT perm, float* PSB_RESTRICT dst, long long npix, int T) {
    const int px = tid % PX;
    const int n_bf = T / R;
    for (int q = tid / PX; q < n_bf; q += kThreads / PX) {
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

// ---- host-side planning ------------------------------------------------------------------------------
inline bool factorise(int T, int* fac, int* nfac) {
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

// storage position pos = k1*(T/R1) + k2*(T/(R1 R2)) + ... holds X[k], k = k1 + R1*(k2 + R2*(...)); fftshift: k -> (k + T/2) % T
inline void build_perm(int T, const int* fac, int nfac, int* perm) {
    for (int pos = 0; pos < T; ++pos) {
        int rem = pos, B = T, k = 0, mult = 1;
        for (int i = 0; i < nfac; ++i) {
            const int sub = B / fac[i];
            const int d = rem / sub;
            rem -= d * sub;
            k += d * mult;
            mult *= fac[i];
            B = sub;
        }
        perm[pos] = (k + T / 2) % T;
    }
}

// pixels per tile: as many as keep the tile within `tile_kb` of shared memory (default 96 KB: two CTAs per SM; the
// environment variable PSB_TACAW_TILE_KB overrides it for tuning runs), else 8 or 4 with the SM to itself
inline int pick_px(int T) {
    static const size_t tile_kb = [] {
        const char* e = std::getenv("PSB_TACAW_TILE_KB");
        const long v = e ? std::atol(e) : 0;
        return (size_t)(v >= 8 && v <= 96 ? v : 96);
    }();
    for (int px : {64, 32, 16, 8})
        if ((size_t)T * px * sizeof(float2) <= (tile_kb << 10)) return px;
    for (int px : {64, 32, 16, 8})
        if ((size_t)T * px * sizeof(float2) <= (96u << 10)) return px;
    if ((size_t)T * 8 * sizeof(float2) <= (200u << 10)) return 8;
    if ((size_t)T * 4 * sizeof(float2) <= (200u << 10)) return 4;
    return 0;
}
inline bool whole_sm(int T) { return (size_t)T * 8 * sizeof(float2) > (96u << 10); }      // one tile per SM: 1024 threads

}  // namespace tw
}  // namespace psb
