// Register-resident Stockham FFT building blocks (complex64, power-of-two line length N).
//
// Layout ("strided register layout"): a line of N points is owned by T = N/E threads; thread j
// holds positions p = j + e*T (e = 0..E-1) in registers v[e].  This is also how a line is read
// from / written to global memory, so for a fixed e consecutive threads touch consecutive
// addresses (coalesced) and no shared-memory staging is needed at the global interface.
//
// Stage k (radix R, Ns = product of earlier radices), thread j, butterfly m (< E/R):
//     b = j + m*T, inputs v[m + t*(E/R)] (= positions b + t*N/R), twiddle w_{Ns*R}^{(b mod Ns)*t},
//     outputs to positions (b div Ns)*Ns*R + (b mod Ns) + u*Ns.
// Every stage reads the same register layout; all but the last exchange through shared memory, and
// the last stage's outputs fall back into the thread's own registers in natural order (index math
// modelled and checked in tools/proto_stockham.py).  A forward transform followed by a pointwise
// multiply and an inverse transform therefore costs only the stage-to-stage exchanges.
//
// Non-power-of-two lengths n use Bluestein's chirp-z identity on top of the same machinery with
// N >= 2n-1 (tables: chirp[p] = exp(-i*pi*p^2/n), Bhat = FFT_N(wrapped conj chirp)/N).
#pragma once
#include "psb_common.cuh"

namespace psb {

// largest supported power-of-two line size
constexpr int kMaxLine = 8192;

constexpr float kC1 = 0.92387953251128674f;   // cos(pi/8)
constexpr float kS1 = 0.38268343236508977f;   // sin(pi/8)
constexpr float kR2 = 0.70710678118654752f;   // sqrt(1/2)

// multiply by w16^K = exp(DIR * 2*pi*i*K/16), K in [0,16), constants folded at compile time
template <int K, int DIR>
PSB_HD float2 mul_w16(float2 a) {
    constexpr int k = ((K % 16) + 16) % 16;
    if constexpr (k == 0) return a;
    else if constexpr (k == 8) return make_float2(-a.x, -a.y);
    else if constexpr (k == 4) return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);      // -/+ i
    else if constexpr (k == 12) return DIR < 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
    else {
        // general: w = (c, DIR*s) with c = cos(2*pi*k/16), s = sin(2*pi*k/16)
        constexpr float c = (k == 1 || k == 15) ? kC1 : (k == 2 || k == 14) ? kR2 : (k == 3 || k == 13) ? kS1
                          : (k == 5 || k == 11) ? -kS1 : (k == 6 || k == 10) ? -kR2 : -kC1;   // k = 7, 9
        constexpr float sabs = (k == 1 || k == 7 || k == 9 || k == 15) ? kS1
                             : (k == 2 || k == 6 || k == 10 || k == 14) ? kR2 : kC1;
        constexpr float s = (k < 8 ? sabs : -sabs) * (DIR < 0 ? -1.0f : 1.0f);
        return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
    }
}

// ---- in-register DFTs, natural-order in and out; DIR = -1 forward, +1 inverse (unnormalised) ----
template <int R, int DIR> struct Dft;

template <int DIR> struct Dft<2, DIR> {
    static PSB_HD void run(float2 (&v)[2]) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <int DIR> struct Dft<4, DIR> {
    static PSB_HD void run(float2 (&v)[4]) {
        float2 s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
        float2 s13 = cadd(v[1], v[3]), d13 = csub(v[1], v[3]);
        v[0] = cadd(s02, s13);
        v[2] = csub(s02, s13);
        // forward: X1 = d02 - i*d13, X3 = d02 + i*d13 ; inverse swaps them
        float2 m = make_float2(d02.x + d13.y, d02.y - d13.x);   // d02 - i*d13
        float2 q = make_float2(d02.x - d13.y, d02.y + d13.x);   // d02 + i*d13
        v[1] = DIR < 0 ? m : q;
        v[3] = DIR < 0 ? q : m;
    }
};

// Cooley-Tukey N = N1*N2: n = N2*n1 + n2, k = k1 + N1*k2
template <int N1, int N2, int DIR>
PSB_HD void dft_ct(float2 (&v)[N1 * N2]);

template <int DIR> struct Dft<8, DIR> {
    static PSB_HD void run(float2 (&v)[8]) {
        float2 a0[4] = {v[0], v[2], v[4], v[6]};
        float2 a1[4] = {v[1], v[3], v[5], v[7]};
        Dft<4, DIR>::run(a0);
        Dft<4, DIR>::run(a1);
        a1[1] = mul_w16<2, DIR>(a1[1]);
        a1[2] = mul_w16<4, DIR>(a1[2]);
        a1[3] = mul_w16<6, DIR>(a1[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = cadd(a0[k], a1[k]);
            v[k + 4] = csub(a0[k], a1[k]);
        }
    }
};

template <int DIR> struct Dft<16, DIR> {
    static PSB_HD void run(float2 (&v)[16]) {
        float2 a[4][4];   // a[n2][k1]
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            float2 t[4] = {v[n2], v[4 + n2], v[8 + n2], v[12 + n2]};
            Dft<4, DIR>::run(t);
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) a[n2][k1] = t[k1];
        }
        a[1][1] = mul_w16<1, DIR>(a[1][1]); a[1][2] = mul_w16<2, DIR>(a[1][2]); a[1][3] = mul_w16<3, DIR>(a[1][3]);
        a[2][1] = mul_w16<2, DIR>(a[2][1]); a[2][2] = mul_w16<4, DIR>(a[2][2]); a[2][3] = mul_w16<6, DIR>(a[2][3]);
        a[3][1] = mul_w16<3, DIR>(a[3][1]); a[3][2] = mul_w16<6, DIR>(a[3][2]); a[3][3] = mul_w16<9, DIR>(a[3][3]);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            float2 t[4] = {a[0][k1], a[1][k1], a[2][k1], a[3][k1]};
            Dft<4, DIR>::run(t);
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = t[k2];
        }
    }
};

// ---- plans: radices per stage for (N, E) -------------------------------------------------------
template <int N, int E> struct Plan;
#define PSB_PLAN(N_, E_, S_, ...)                                  \
    template <> struct Plan<N_, E_> {                               \
        static constexpr int S = S_;                                \
        static constexpr int R[4] = {__VA_ARGS__};                  \
    };
PSB_PLAN(16, 8, 2, 2, 8, 1, 1)
PSB_PLAN(32, 8, 2, 4, 8, 1, 1)
PSB_PLAN(64, 8, 2, 8, 8, 1, 1)
PSB_PLAN(128, 8, 3, 8, 2, 8, 1)
PSB_PLAN(256, 8, 3, 8, 4, 8, 1)
PSB_PLAN(512, 8, 3, 8, 8, 8, 1)
PSB_PLAN(256, 16, 2, 16, 16, 1, 1)
PSB_PLAN(512, 16, 3, 16, 2, 16, 1)
PSB_PLAN(1024, 16, 3, 16, 4, 16, 1)
PSB_PLAN(2048, 16, 3, 16, 8, 16, 1)
PSB_PLAN(4096, 16, 3, 16, 16, 16, 1)
PSB_PLAN(8192, 16, 4, 16, 2, 16, 16)
#undef PSB_PLAN

// Shared-memory addressing of position q of line c inside a tile of W lines.
//   COLS (lines strided in global memory, lanes run across lines):  q*W + c      (conflict-free)
//   ROWS (lines contiguous, lanes run along the line):              c*NP + q + q/16, NP = N + N/16
template <int N, int W, bool COLS>
struct SmemMap {
    static constexpr int NP = N + N / 16;
    static constexpr size_t kBytes = COLS ? (size_t)N * W * sizeof(float2) : (size_t)NP * W * sizeof(float2);
    static PSB_HD int at(int c, int q) { return COLS ? q * W + c : c * NP + q + (q >> 4); }
};

// elements per thread used for a given line size (host and device agree through this one rule)
constexpr int elems_for(int N) { return N <= 128 ? 8 : 16; }

// Twiddle table layout ("staged"): for every stage with NS > 1 a block of (R-1)*NS entries,
//     block[(t-1)*NS + k] = exp(-2*pi*i*k*t/(NS*R)),  t = 1..R-1, k = 0..NS-1,
// blocks concatenated in stage order.  Threads with consecutive k read consecutive entries, so the
// loads coalesce (the flat exp(-2*pi*i*k/N) table made row-oriented passes fetch up to 15 cache
// lines per warp request -- ncu profile r1 v1).
template <int N, int E>
constexpr int twiddle_offset(int stage) {
    int off = 0, ns = 1;
    for (int s = 0; s < stage; ++s) {
        if (ns > 1) off += (Plan<N, E>::R[s] - 1) * ns;
        ns *= Plan<N, E>::R[s];
    }
    return off;
}
template <int N, int E>
constexpr int twiddle_count() { return twiddle_offset<N, E>(Plan<N, E>::S); }

// One Stockham stage.  `tw` points at this stage's block of the staged twiddle table.
template <int N, int E, int W, bool COLS, int R, int NS, int DIR, bool LAST, class Ctx>
PSB_D void fft_stage(const Ctx& cx, float2 (&v)[E], float2* sm, int c, int j, const float2* PSB_RESTRICT tw) {
    constexpr int T = N / E;
    constexpr int NB = E / R;
    using Map = SmemMap<N, W, COLS>;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        float2 a[R];
#pragma unroll
        for (int t = 0; t < R; ++t) a[t] = v[m + t * NB];
        const int b = j + m * T;
        if constexpr (NS > 1) {
            const int k = b & (NS - 1);
#pragma unroll
            for (int t = 1; t < R; ++t) {
                float2 w = __ldg(&tw[(t - 1) * NS + k]);
                a[t] = DIR < 0 ? cmul(a[t], w) : cmulc(a[t], w);
            }
        }
        Dft<R, DIR>::run(a);
        if constexpr (LAST) {
#pragma unroll
            for (int u = 0; u < R; ++u) v[m + u * NB] = a[u];
        } else {
            const int q0 = (b / NS) * (NS * R) + (b & (NS - 1));
#pragma unroll
            for (int u = 0; u < R; ++u) sm[Map::at(c, q0 + u * NS)] = a[u];
        }
    }
    if constexpr (!LAST) {
        cx.sync();
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = sm[Map::at(c, j + e * T)];
        cx.sync();
    }
}

// Full line FFT (unnormalised).  c = line within the tile, j = thread within the line.
template <int N, int E, int W, bool COLS, int DIR, class Ctx>
PSB_D void fft_line(const Ctx& cx, float2 (&v)[E], float2* sm, int c, int j, const float2* PSB_RESTRICT tw) {
    using P = Plan<N, E>;
    constexpr int R0 = P::R[0], R1 = P::R[1], R2 = P::R[2];
    fft_stage<N, E, W, COLS, R0, 1, DIR, P::S == 1>(cx, v, sm, c, j, tw);
    if constexpr (P::S >= 2) fft_stage<N, E, W, COLS, R1, R0, DIR, P::S == 2>(cx, v, sm, c, j, tw + twiddle_offset<N, E>(1));
    if constexpr (P::S >= 3) fft_stage<N, E, W, COLS, R2, R0 * R1, DIR, P::S == 3>(cx, v, sm, c, j, tw + twiddle_offset<N, E>(2));
    if constexpr (P::S >= 4) fft_stage<N, E, W, COLS, P::R[3], R0 * R1 * R2, DIR, true>(cx, v, sm, c, j, tw + twiddle_offset<N, E>(3));
}

// Tables one line transform needs.  For the direct (power-of-two) path only `tw` is used.
struct FftTables {
    const float2* tw;      // staged twiddle table of the (N, E) plan, see twiddle_offset
    const float2* chirp;   // [n]  exp(-i*pi*p^2/n)            (Bluestein only)
    const float2* bhat;    // [N]  FFT_N(wrapped conj chirp)/N  (Bluestein only)
    int n;                 // logical line length (== N on the direct path)
};

// DFT of logical length n (DIR = -1 forward, +1 unnormalised inverse) in the strided register
// layout of an N-point line.  BLUE=false requires n == N.  With BLUE=true, on entry registers with
// p >= n are ignored, on exit they are zero.
template <int N, int E, int W, bool COLS, int DIR, bool BLUE, class Ctx>
PSB_D void dft_line(const Ctx& cx, float2 (&v)[E], float2* sm, int c, int j, const FftTables& tb) {
    if constexpr (!BLUE) {
        fft_line<N, E, W, COLS, DIR>(cx, v, sm, c, j, tb.tw);
    } else {
        constexpr int T = N / E;
        // inverse via conj(F(conj x)); forward: X[k] = chirp[k] * sum_p x[p] chirp[p] conj(chirp[k-p])
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int p = j + e * T;
            if (p < tb.n) {
                float2 x = DIR < 0 ? v[e] : cconj(v[e]);
                v[e] = cmul(x, __ldg(&tb.chirp[p]));
            } else {
                v[e] = make_float2(0.f, 0.f);
            }
        }
        fft_line<N, E, W, COLS, -1>(cx, v, sm, c, j, tb.tw);
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cmul(v[e], __ldg(&tb.bhat[j + e * T]));
        fft_line<N, E, W, COLS, +1>(cx, v, sm, c, j, tb.tw);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int p = j + e * T;
            if (p < tb.n) {
                float2 y = cmul(v[e], __ldg(&tb.chirp[p]));
                v[e] = DIR < 0 ? y : cconj(y);
            } else {
                v[e] = make_float2(0.f, 0.f);
            }
        }
    }
}

}  // namespace psb
