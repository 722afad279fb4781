// Register butterflies, stages and host-side planning of the tiled mixed-radix TACAW transform (tacaw_fast.cu).
// Device code in the product build; the same source compiles for the host (PSB_EMU) so tests/emu/tacaw_fast_harness.cpp
// can run a tile's stage sequence thread by thread and check index arithmetic, twiddle selection and the output
// permutation against numpy without a GPU.
#pragma once
#include "fast_fft.cuh"
#include "psb_common.cuh"

#include <cstdlib>

namespace psb {
namespace tw {

constexpr int kMaxFactors = 16;

using fast::cpx;

// forward R-point DFTs in registers (W_R = exp(-2 pi i / R)) on packed complex values: every complex add is one FADD2,
// every real-constant scaling one FMUL2 / FFMA2, the (y, -x) rotations ride on the operand swizzle of the packed
// instructions.  (The scalar float2 version of round 1 issued 77 instructions per element at T = 100 and the kernel was
// issue-bound: 70 % of the issue slots busy at 0.61 of the copy bandwidth, ncu r2v.)
PSB_D cpx both(float v) { return fast::c_make(v, v); }
PSB_D void dft2(cpx* x) {
    const cpx a = x[0], b = x[1];
    x[0] = fast::add2(a, b);
    x[1] = fast::sub2(a, b);
}
PSB_D void dft3(cpx* x) {
    const float c = -0.5f, s = 0.86602540378443865f;
    const cpx t1 = fast::add2(x[1], x[2]), t2 = fast::sub2(x[1], x[2]);
    const cpx m = fast::fma2(t1, both(c), x[0]);
    const cpx u = fast::mul2(t2, both(s));
    x[0] = fast::add2(x[0], t1);
    x[1] = fast::sub_ib(m, u);          // m - i*s*(x1 - x2)
    x[2] = fast::add_ib(m, u);
}
PSB_D void dft4(cpx* x) { fast::radix4<-1>(x[0], x[1], x[2], x[3]); }
PSB_D void dft5(cpx* x) {
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;      // cos(2 pi/5), cos(4 pi/5)
    const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;       // sin(2 pi/5), sin(4 pi/5)
    const cpx a1 = fast::add2(x[1], x[4]), b1 = fast::sub2(x[1], x[4]);
    const cpx a2 = fast::add2(x[2], x[3]), b2 = fast::sub2(x[2], x[3]);
    const cpx m1 = fast::fma2(a2, both(c2), fast::fma2(a1, both(c1), x[0]));
    const cpx m2 = fast::fma2(a2, both(c1), fast::fma2(a1, both(c2), x[0]));
    const cpx u1 = fast::fma2(b2, both(s2), fast::mul2(b1, both(s1)));      // s1 b1 + s2 b2
    const cpx u2 = fast::fma2(b2, both(-s1), fast::mul2(b1, both(s2)));     // s2 b1 - s1 b2
    x[0] = fast::add2(x[0], fast::add2(a1, a2));
    x[1] = fast::sub_ib(m1, u1);
    x[4] = fast::add_ib(m1, u1);
    x[2] = fast::sub_ib(m2, u2);
    x[3] = fast::add_ib(m2, u2);
}

template <int R>
PSB_D void dft(cpx* x);

// N1 x N2 points with coprime factors by the prime-factor (Good-Thomas) mapping: input n = N2 n1 + N1 n2, output
// k = I1 k1 + I2 k2 (mod N) with I1 = N2 (N2^-1 mod N1), I2 = N1 (N1^-1 mod N2) -- no twiddles between the two layers, and
// with everything unrolled the index maps are register renaming.  A stage of radix 10 or 20 replaces two stages of
// radix 2/4 and 5: one trip through shared memory, one barrier and one set of twiddle multiplies less per element.
template <int N1, int N2, int I1, int I2>
PSB_D void dft_pfa(cpx* x) {
    constexpr int N = N1 * N2;
    cpx y[N2][N1];
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) {
#pragma unroll
        for (int n1 = 0; n1 < N1; ++n1) y[n2][n1] = x[(N2 * n1 + N1 * n2) % N];
        dft<N1>(y[n2]);
    }
#pragma unroll
    for (int k1 = 0; k1 < N1; ++k1) {
        cpx z[N2];
#pragma unroll
        for (int n2 = 0; n2 < N2; ++n2) z[n2] = y[n2][k1];
        dft<N2>(z);
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) x[(I1 * k1 + I2 * k2) % N] = z[k2];
    }
}

template <int R>
PSB_D void dft(cpx* x) {
    if constexpr (R == 2) dft2(x);
    else if constexpr (R == 3) dft3(x);
    else if constexpr (R == 4) dft4(x);
    else if constexpr (R == 5) dft5(x);
    else if constexpr (R == 10) dft_pfa<2, 5, 5, 6>(x);
    else dft_pfa<4, 5, 5, 16>(x);
    static_assert(R == 2 || R == 3 || R == 4 || R == 5 || R == 10 || R == 20, "radix");
}

// Stage s of the decimation in frequency splits every block of B elements into R sub-blocks of B / R:
//     y[k1 * sub + n] = DFT_R( x[. * sub + n] )[k1] * W_B^(n k1),   W_B^m = tw[m * T / B].
// The first stage (B = T) takes its inputs straight from global memory, the last one (sub = 1, no twiddles) hands
// |.|^2 straight to global memory, so an element crosses shared memory twice per middle stage and once at each end.

template <int R, int PX, int kThreads>
PSB_D void first_stage(unsigned tid, float2* data_, const float2* PSB_RESTRICT src_, long long stride_frame, bool live,
                       const float2* PSB_RESTRICT tw_, int T) {
    cpx* data = reinterpret_cast<cpx*>(data_);
    const cpx* PSB_RESTRICT src = reinterpret_cast<const cpx*>(src_);
    const cpx* PSB_RESTRICT tw = reinterpret_cast<const cpx*>(tw_);
    const cpx zero = fast::c_make(0.f, 0.f);
    const int sub = T / R;
    const int px = tid % PX;
    constexpr int step = kThreads / PX;
    // Two butterflies per trip: the 2*R global loads are all issued before the first butterfly is computed.  With one
    // butterfly per trip a CTA kept R*kThreads*8 bytes in flight -- about half of what the HBM latency-bandwidth product
    // asks of an SM (ncu r1j: long-scoreboard stalls 4.0 per issue, 0.61 of the copy bandwidth).
    int n = tid / PX;
    if constexpr (R <= 5)          // a radix-10 / 20 butterfly has that many loads in flight by itself
    for (; n + step < sub; n += 2 * step) {
        cpx x[R], y[R];
#pragma unroll
        for (int i = 0; i < R; ++i) x[i] = live ? src[(long long)(n + i * sub) * stride_frame] : zero;
#pragma unroll
        for (int i = 0; i < R; ++i) y[i] = live ? src[(long long)(n + step + i * sub) * stride_frame] : zero;
        dft<R>(x);
        data[n * PX + px] = x[0];
#pragma unroll
        for (int k = 1; k < R; ++k) data[(k * sub + n) * PX + px] = fast::cmulp(x[k], __ldg(&tw[n * k]));
        dft<R>(y);
        data[(n + step) * PX + px] = y[0];
#pragma unroll
        for (int k = 1; k < R; ++k) data[(k * sub + n + step) * PX + px] = fast::cmulp(y[k], __ldg(&tw[(n + step) * k]));
    }
    for (; n < sub; n += step) {
        cpx x[R];
#pragma unroll
        for (int i = 0; i < R; ++i) x[i] = live ? src[(long long)(n + i * sub) * stride_frame] : zero;
        dft<R>(x);
        data[n * PX + px] = x[0];
#pragma unroll
        for (int k = 1; k < R; ++k) data[(k * sub + n) * PX + px] = fast::cmulp(x[k], __ldg(&tw[n * k]));
    }
}

template <int R, int PX, int kThreads>
PSB_D void mid_stage(unsigned tid, float2* data_, const float2* PSB_RESTRICT tw_, int T, int B) {
    cpx* data = reinterpret_cast<cpx*>(data_);
    const cpx* PSB_RESTRICT tw = reinterpret_cast<const cpx*>(tw_);
    const int sub = B / R;
    const int tstep = T / B;
    const int px = tid % PX;
    const int n_bf = T / R;
    constexpr int step = kThreads / PX;
    // butterfly bi = q * sub + n; (q, n) advance with the loop instead of a division per butterfly
    int bi = tid / PX;
    int q = bi / sub, n = bi - q * sub;
    const int dq = step / sub, dn = step - dq * sub;
    for (; bi < n_bf; bi += step) {
        cpx* base = data + (q * B + n) * PX + px;
        cpx x[R];
#pragma unroll
        for (int i = 0; i < R; ++i) x[i] = base[i * sub * PX];
        dft<R>(x);
        base[0] = x[0];
        const int tn = n * tstep;
#pragma unroll
        for (int k = 1; k < R; ++k) base[k * sub * PX] = fast::cmulp(x[k], __ldg(&tw[tn * k]));
        n += dn;
        q += dq;
        if (n >= sub) {
            n -= sub;
            ++q;
        }
    }
}

// kFromGlobal: the transform has a single stage (T = R), inputs come from global memory as well
template <int R, int PX, int kThreads, bool kFromGlobal>
PSB_D void last_stage(unsigned tid, const float2* data_, const float2* PSB_RESTRICT src_, long long stride_frame, bool live,
                      const int* PSB_RESTRICT perm, float* PSB_RESTRICT dst, long long npix, int T) {
    const cpx* data = reinterpret_cast<const cpx*>(data_);
    const cpx* PSB_RESTRICT src = reinterpret_cast<const cpx*>(src_);
    const int px = tid % PX;
    const int n_bf = T / R;
    for (int q = tid / PX; q < n_bf; q += kThreads / PX) {
        cpx x[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if (kFromGlobal) x[i] = live ? src[(long long)i * stride_frame] : fast::c_make(0.f, 0.f);
            else x[i] = data[(q * R + i) * PX + px];
        }
        dft<R>(x);
        if (!live) continue;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int pos = q * R + k;
            // position 0 holds X[0] = T * mean: psi - <psi>_t differs from psi in this bin only (and is 0 there)
            const cpx sq = fast::mul2(x[k], x[k]);
            const float v = pos == 0 ? 0.f : fast::c_re(sq) + fast::c_im(sq);
            dst[(long long)__ldg(&perm[pos]) * npix] = v;
        }
    }
}

// BIG: 0 = the kernel instance with radices up to 5, 1 = also radix 10, 2 = also radix 20 (more registers each; separate
// instances so that frame counts without such factors keep their occupancy)
#define PSB_TW_DISPATCH(R_, CALL)                 \
    do {                                          \
        if (BIG >= 2 && (R_) == 20) { constexpr int R = BIG >= 2 ? 20 : 2; CALL; }      \
        else if (BIG >= 1 && (R_) == 10) { constexpr int R = BIG >= 1 ? 10 : 2; CALL; } \
        else if ((R_) == 5) { constexpr int R = 5; CALL; }      \
        else if ((R_) == 4) { constexpr int R = 4; CALL; } \
        else if ((R_) == 3) { constexpr int R = 3; CALL; } \
        else { constexpr int R = 2; CALL; }                \
    } while (0)

// ---- host-side planning ------------------------------------------------------------------------------
inline bool factorise_with(int T, const int* radices, int nr, int* fac, int* nfac) {
    int n = 0, r = T;
    for (int i = 0; i < nr; ++i)
        while (r % radices[i] == 0 && r > 1) {
            if (n == kMaxFactors) return false;
            fac[n++] = radices[i];
            r /= radices[i];
        }
    *nfac = n;
    return r == 1 && n > 0;
}
// fewest stages first; among equals the plan without radix 20 (fewer registers, better balanced stages):
// 20 -> {20}, 100 -> {10, 10}, 500 -> {10, 10, 5}, 2000 -> {20, 10, 10}
inline bool factorise(int T, int* fac, int* nfac) {
    const int r20[6] = {20, 10, 5, 4, 3, 2}, r10[5] = {10, 5, 4, 3, 2};
    int f10[kMaxFactors], n10 = 0;
    if (!factorise_with(T, r10, 5, f10, &n10)) return false;
    if (factorise_with(T, r20, 6, fac, nfac) && *nfac < n10) {
        // one radix 20 is enough when the rest folds into tens: {20, 20, 5} -> {20, 10, 10}
        for (int i = 0; i + 2 < *nfac; ++i)
            if (fac[i] == 20 && fac[i + 1] == 20 && fac[*nfac - 1] == 5) {
                fac[i + 1] = 10;
                fac[*nfac - 1] = 10;
                break;
            }
        return true;
    }
    for (int i = 0; i < n10; ++i) fac[i] = f10[i];
    *nfac = n10;
    return true;
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
