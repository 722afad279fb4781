// Line FFTs of the fused kernels (fast_path.cu, sf_nufft.cu): N = 256, 512, 1024 or 2048 complex64 points held by
// T = N/16 threads, 16 points per thread in the strided register layout of fft_core.cuh
// (thread j owns positions j + e*T), built for Blackwell's packed fp32 pipe:
//
//   * every complex add/sub is one FADD2, every complex multiply one FMUL2 + one FFMA2 (the operand
//     swizzle / per-half sign modifiers of the sm_100 F32x2 instructions absorb the (-y, x) rotation),
//     so a radix-16 butterfly is 81 issue slots instead of ~165 scalar ones;
//   * the inter-stage twiddles of a thread depend only on its position j inside the line, not on the
//     tile: a persistent kernel loads them once into registers (`Twiddles`) and reuses them for every
//     tile and both directions (the inverse multiplies by the conjugate);
//   * stages exchange through shared memory behind a policy object (`Xchg`) that knows the layout and
//     the synchronisation scope: a warp-private buffer + __syncwarp for contiguous lines (16 or 32
//     threads of one warp own a line), CTA barriers + two alternating buffers for strided lines.
//
// The same source compiles for the host (scalar arithmetic, PSB_EMU) so tests/ can check the index
// arithmetic of every stage against numpy without a GPU (tests/emu/fast_fft_harness.cpp).
#pragma once
#include "fft_core.cuh"

namespace psb {
namespace fast {

// ---- packed complex arithmetic ----------------------------------------------------------------
// On the device a complex number lives in ONE 64-bit register pair (`cpx` = b64) from load to store, so
// ptxas never has to re-pair the halves (a float2 round trip through st.v2.f32 serialised every store
// through one register pair -- ncu r1d); on the host (tests) it is a plain float2.
#if defined(__CUDA_ARCH__)
typedef unsigned long long cpx;
PSB_D cpx c_make(float x, float y) {
    cpx d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y));
    return d;
}
PSB_D float c_re(cpx a) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a));
    return x;
}
PSB_D float c_im(cpx a) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a));
    return y;
}
PSB_D cpx add2(cpx a, cpx b) {
    cpx d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
PSB_D cpx sub2(cpx a, cpx b) {
    cpx d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
PSB_D cpx mul2(cpx a, cpx b) {
    cpx d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
PSB_D cpx fma2(cpx a, cpx b, cpx c) {
    cpx d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
#else
typedef float2 cpx;
PSB_D cpx c_make(float x, float y) { return make_float2(x, y); }
PSB_D float c_re(cpx a) { return a.x; }
PSB_D float c_im(cpx a) { return a.y; }
PSB_D cpx add2(cpx a, cpx b) { return make_float2(a.x + b.x, a.y + b.y); }
PSB_D cpx sub2(cpx a, cpx b) { return make_float2(a.x - b.x, a.y - b.y); }
PSB_D cpx mul2(cpx a, cpx b) { return make_float2(a.x * b.x, a.y * b.y); }
PSB_D cpx fma2(cpx a, cpx b, cpx c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif
static_assert(sizeof(cpx) == sizeof(float2), "cpx is a bit-cast of float2");

// a * w
PSB_D cpx cmulp(cpx a, cpx w) {
    return fma2(c_make(-c_im(a), c_re(a)), c_make(c_im(w), c_im(w)), mul2(a, c_make(c_re(w), c_re(w))));
}
// a * conj(w)
PSB_D cpx cmulcp(cpx a, cpx w) {
    return fma2(c_make(c_im(a), -c_re(a)), c_make(c_im(w), c_im(w)), mul2(a, c_make(c_re(w), c_re(w))));
}
// a - i*b  and  a + i*b
PSB_D cpx sub_ib(cpx a, cpx b) { return fma2(c_make(c_im(b), c_re(b)), c_make(1.f, -1.f), a); }
PSB_D cpx add_ib(cpx a, cpx b) { return fma2(c_make(c_im(b), c_re(b)), c_make(-1.f, 1.f), a); }

// multiply by exp(DIR * 2*pi*i*K/16), constants folded
template <int K, int DIR>
PSB_D cpx rot16(cpx a) {
    constexpr int k = ((K % 16) + 16) % 16;
    if constexpr (k == 0) return a;
    else if constexpr (k == 8) return c_make(-c_re(a), -c_im(a));
    else if constexpr (k == 4) return DIR < 0 ? mul2(c_make(c_im(a), c_re(a)), c_make(1.f, -1.f))     // -i*a
                                              : mul2(c_make(c_im(a), c_re(a)), c_make(-1.f, 1.f));    // +i*a
    else if constexpr (k == 12) return DIR < 0 ? mul2(c_make(c_im(a), c_re(a)), c_make(-1.f, 1.f))
                                               : mul2(c_make(c_im(a), c_re(a)), c_make(1.f, -1.f));
    else {
        constexpr float c = (k == 1 || k == 15) ? kC1 : (k == 2 || k == 14) ? kR2 : (k == 3 || k == 13) ? kS1
                          : (k == 5 || k == 11) ? -kS1 : (k == 6 || k == 10) ? -kR2 : -kC1;
        constexpr float sabs = (k == 1 || k == 7 || k == 9 || k == 15) ? kS1
                             : (k == 2 || k == 6 || k == 10 || k == 14) ? kR2 : kC1;
        constexpr float s = (k < 8 ? sabs : -sabs) * (DIR < 0 ? -1.0f : 1.0f);
        return cmulp(a, c_make(c, s));
    }
}

template <int DIR>
PSB_D void radix4(cpx& a0, cpx& a1, cpx& a2, cpx& a3) {
    const cpx s02 = add2(a0, a2), d02 = sub2(a0, a2);
    const cpx s13 = add2(a1, a3), d13 = sub2(a1, a3);
    a0 = add2(s02, s13);
    a2 = sub2(s02, s13);
    const cpx m = sub_ib(d02, d13), q = add_ib(d02, d13);   // forward: X1 = d02 - i*d13, X3 = d02 + i*d13
    a1 = DIR < 0 ? m : q;
    a3 = DIR < 0 ? q : m;
}

template <int DIR>
PSB_D void radix2(cpx& a0, cpx& a1) {
    const cpx s = add2(a0, a1), d = sub2(a0, a1);
    a0 = s;
    a1 = d;
}

// 8-point DFT in place (2 x 4 Cooley-Tukey: n = 2*n1 + n2, k = k1 + 4*k2), natural order in and out
template <int DIR>
PSB_D void radix8(cpx (&a)[8]) {
    cpx e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];      // n2 = 0
    cpx o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];      // n2 = 1
    radix4<DIR>(e0, e1, e2, e3);
    radix4<DIR>(o0, o1, o2, o3);
    o1 = rot16<2, DIR>(o1);                               // W8^1
    o2 = rot16<4, DIR>(o2);                               // W8^2
    o3 = rot16<6, DIR>(o3);                               // W8^3
    a[0] = add2(e0, o0); a[4] = sub2(e0, o0);
    a[1] = add2(e1, o1); a[5] = sub2(e1, o1);
    a[2] = add2(e2, o2); a[6] = sub2(e2, o2);
    a[3] = add2(e3, o3); a[7] = sub2(e3, o3);
}

// 16-point DFT (4 x 4 Cooley-Tukey: n = 4*n1 + n2, k = k1 + 4*k2) as a stream: `in(idx)` produces input idx when
// the first butterfly layer needs it, `out(idx, value)` takes output idx as soon as the second layer has it.
// Fusing the loads / pointwise multiplies / stores of a pass into these functors keeps only a few of the 16
// values in flight around each memory operation; as separate 16-wide loops they made ptxas funnel every store
// through one register pair (two MOVs per store and a serialised chain, ncu r1d/r1f).
template <int DIR, class In, class Out>
PSB_D void radix16_io(const In& in, const Out& out) {
    cpx a[4][4];   // a[n2][k1]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        cpx t0 = in(n2), t1 = in(4 + n2), t2 = in(8 + n2), t3 = in(12 + n2);
        radix4<DIR>(t0, t1, t2, t3);
        a[n2][0] = t0; a[n2][1] = t1; a[n2][2] = t2; a[n2][3] = t3;
    }
    a[1][1] = rot16<1, DIR>(a[1][1]); a[1][2] = rot16<2, DIR>(a[1][2]); a[1][3] = rot16<3, DIR>(a[1][3]);
    a[2][1] = rot16<2, DIR>(a[2][1]); a[2][2] = rot16<4, DIR>(a[2][2]); a[2][3] = rot16<6, DIR>(a[2][3]);
    a[3][1] = rot16<3, DIR>(a[3][1]); a[3][2] = rot16<6, DIR>(a[3][2]); a[3][3] = rot16<9, DIR>(a[3][3]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        cpx t0 = a[0][k1], t1 = a[1][k1], t2 = a[2][k1], t3 = a[3][k1];
        radix4<DIR>(t0, t1, t2, t3);
        out(k1, t0); out(k1 + 4, t1); out(k1 + 8, t2); out(k1 + 12, t3);
    }
}

// ---- per-thread persistent twiddles ---------------------------------------------------------------
// N = 256  (radix 16 x 16):      w[t-1] = exp(-2*pi*i*j*t/256),  t = 1..15
// N = 512  (radix 16 x 2 x 16):  w[t-1] = exp(-2*pi*i*j*t/512),  w2 = exp(-2*pi*i*(j & 15)/32)
//                                (the radix-2 stage is folded into the last stage's loads, see line_fft)
// N = 1024 (radix 16 x 4 x 16):  w[t-1] = exp(-2*pi*i*j*t/1024), w4[t-1] = exp(-2*pi*i*(j & 15)*t/64), t = 1..3
//                                (the radix-4 stage runs in place in the exchange buffer, see line_fft)
// N = 2048 (radix 16 x 8 x 16):  w[t-1] = exp(-2*pi*i*j*t/2048), w4[t-1] = exp(-2*pi*i*(j & 15)*t/128), t = 1..7
template <int N>
struct Twiddles {
    cpx w[15];
    cpx w2;
    cpx w4[N == 2048 ? 7 : 3];
    // `staged` is the staged table of Plan<N, 16> built by tables.cu (layout: fft_core.cuh twiddle_offset)
    PSB_D void load(const float2* PSB_RESTRICT staged_f2, int j) {
        const cpx* PSB_RESTRICT staged = reinterpret_cast<const cpx*>(staged_f2);
        static_assert(N == 256 || N == 512 || N == 1024 || N == 2048, "fast path line sizes");
        w2 = c_make(1.f, 0.f);
#pragma unroll
        for (int t = 0; t < (N == 2048 ? 7 : 3); ++t) w4[t] = w2;
        if constexpr (N == 256) {
#pragma unroll
            for (int t = 1; t < 16; ++t) w[t - 1] = staged[(t - 1) * 16 + j];
        } else if constexpr (N == 512) {
            w2 = staged[j & 15];                                           // stage 2 block: R = 2, NS = 16
#pragma unroll
            for (int t = 1; t < 16; ++t) w[t - 1] = staged[16 + (t - 1) * 32 + j];   // stage 3 block: NS = 32
        } else if constexpr (N == 1024) {
#pragma unroll
            for (int t = 1; t < 4; ++t) w4[t - 1] = staged[(t - 1) * 16 + (j & 15)];  // stage 2 block: R = 4, NS = 16
#pragma unroll
            for (int t = 1; t < 16; ++t) w[t - 1] = staged[48 + (t - 1) * 64 + j];   // stage 3 block: NS = 64
        } else {
#pragma unroll
            for (int t = 1; t < 8; ++t) w4[t - 1] = staged[(t - 1) * 16 + (j & 15)];  // stage 2 block: R = 8, NS = 16
#pragma unroll
            for (int t = 1; t < 16; ++t) w[t - 1] = staged[112 + (t - 1) * 128 + j]; // stage 3 block: NS = 128
        }
    }
};


// ---- the line transform -----------------------------------------------------------------------------
// Xchg policy:  cpx* buf(int i)         exchange buffer of the i-th exchange of the current tile
//               int at(int q)           element index of position q of this thread's line in that buffer
//               void after_store(int i) all stores of exchange i visible to the line's threads
//               void after_load(int i)  all loads of exchange i done (buffer reusable)
//               void mid_sync(int i)    N = 1024: the in-place middle stage's stores to buffer i are visible
// in(e)  -> the thread's input at position j + e*T (e = 0..15), called once per e while the first stage runs
// out(e, value) <- the transform at position j + e*T, called once per e while the last stage runs
// `hook()` runs right after the first exchange's stores are visible: by then every thread of the line's
// sync scope has CONSUMED what `in` gave it (the stores depend on it), which is the earliest point at which
// the buffer `in` read from may be handed back to the async proxy.
// `before_last()` runs before the last stage starts calling `out` (e.g. wait for the operand `out` multiplies by).
struct NoHook {
    PSB_D void operator()() const {}
};

template <int N, int DIR, class In, class Out, class Xchg, class Hook = NoHook, class Hook2 = NoHook>
PSB_D void line_fft(const In& in, const Out& out, const Twiddles<N>& tw, int j, const Xchg& x, int xi0,
                    const Hook& hook = Hook(), const Hook2& before_last = Hook2()) {
    constexpr int T = N / 16;
    // stage 1: radix 16 over positions j + t*T, outputs to 16*j + u
    {
        cpx* sm = x.buf(xi0);
        radix16_io<DIR>(in, [&](int u, cpx val) { sm[x.at(16 * j + u)] = val; });
        x.after_store(xi0);
        hook();
    }
    if constexpr (N == 1024) {
        // Middle stage (radix 4, NS = 16) IN PLACE in the exchange buffer.  In the Stockham indexing of fft_core.cuh the
        // thread's butterfly m reads logical positions b + 256*t (b = j + 64*m, twiddle exp(-+2*pi*i*(b & 15)*t/64), and
        // b & 15 = j & 15 for every m) and would write positions (b >> 4)*64 + (b & 15) + 16*u.  Instead output u is
        // written back to the slot input t = u came from, so no thread touches another thread's words between the two
        // barriers and one buffer serves both exchanges; the last stage then fetches logical position j + 64*t from the
        // slot that holds it: thread (j & 15) + 16*(t & 3), butterfly t >> 2, output j >> 4, i.e. slot
        // (j & 15) + 16*(t & 3) + 64*(t >> 2) + 256*(j >> 4)  (checked against numpy on the host, tests/test_fast_fft_host.py).
        cpx* sm = x.buf(xi0);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int b = j + 64 * m;
            cpx a0 = sm[x.at(b)], a1 = sm[x.at(b + 256)], a2 = sm[x.at(b + 512)], a3 = sm[x.at(b + 768)];
            a1 = DIR < 0 ? cmulp(a1, tw.w4[0]) : cmulcp(a1, tw.w4[0]);
            a2 = DIR < 0 ? cmulp(a2, tw.w4[1]) : cmulcp(a2, tw.w4[1]);
            a3 = DIR < 0 ? cmulp(a3, tw.w4[2]) : cmulcp(a3, tw.w4[2]);
            radix4<DIR>(a0, a1, a2, a3);
            sm[x.at(b)] = a0; sm[x.at(b + 256)] = a1; sm[x.at(b + 512)] = a2; sm[x.at(b + 768)] = a3;
        }
        x.mid_sync(xi0);
    }
    if constexpr (N == 2048) {
        // The same with radix 8 (T = 128 threads, two butterflies per thread): butterfly m reads b + 256*t, b = j + 128*m,
        // twiddle exp(-+2*pi*i*(j & 15)*t/128); output u returns to the slot of input u.  The last stage finds logical
        // position j + 128*t in slot (j & 15) + 16*(t & 7) + 128*(t >> 3) + 256*(j >> 4).
        cpx* sm = x.buf(xi0);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int b = j + 128 * m;
            cpx a[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) a[t] = sm[x.at(b + 256 * t)];
#pragma unroll
            for (int t = 1; t < 8; ++t) a[t] = DIR < 0 ? cmulp(a[t], tw.w4[t - 1]) : cmulcp(a[t], tw.w4[t - 1]);
            radix8<DIR>(a);
#pragma unroll
            for (int t = 0; t < 8; ++t) sm[x.at(b + 256 * t)] = a[t];
        }
        x.mid_sync(xi0);
    }
    // last stage: radix 16 with the thread's own twiddles; outputs come out in natural strided order
    const cpx* sl = x.buf(xi0);
    before_last();
    if constexpr (N == 1024 || N == 2048) {
        constexpr int R2 = N / 256;                  // radix of the middle stage
        const int slot0 = (j & 15) + 256 * (j >> 4);
        radix16_io<DIR>(
            [&](int t) {
                const cpx a = sl[x.at(slot0 + 16 * (t % R2) + T * (t / R2))];
                if (t == 0) return a;
                return DIR < 0 ? cmulp(a, tw.w[t > 0 ? t - 1 : 0]) : cmulcp(a, tw.w[t > 0 ? t - 1 : 0]);
            },
            out);
    } else if constexpr (N == 512) {
        // The radix-2 stage (NS = 16: butterfly pairs positions b and b + 256, twiddle exp(-+2*pi*i*(b & 15)/32) = w2)
        // is folded into the loads of the last stage instead of taking its own trip through shared memory: its
        // outputs land at (b/16)*32 + (b & 15) [sum] and + 16 [difference], and the last stage wants positions
        // j + 32*t of that array -- for j < 16 those are the sums of b = 16*t + j, for j >= 16 the differences of
        // b = 16*t + (j - 16).  So thread j loads both butterfly inputs itself and keeps one output; the partner
        // thread j ^ 16 loads the same two words (a broadcast) and keeps the other.  32 loads instead of
        // 16 loads + 16 stores + 16 loads, one exchange and one barrier less per transform, 8 extra complex
        // multiplies per thread.
        const int jj = j & 15;
        const float sg = (j & 16) ? -1.f : 1.f;
        const cpx sg2 = c_make(sg, sg);
        radix16_io<DIR>(
            [&](int t) {
                const int b = 16 * t + jj;
                const cpx lo = sl[x.at(b)], hi = sl[x.at(b + 256)];
                const cpx hw = DIR < 0 ? cmulp(hi, tw.w2) : cmulcp(hi, tw.w2);
                const cpx a = fma2(hw, sg2, lo);
                if (t == 0) return a;
                return DIR < 0 ? cmulp(a, tw.w[t > 0 ? t - 1 : 0]) : cmulcp(a, tw.w[t > 0 ? t - 1 : 0]);
            },
            out);
    } else {
        radix16_io<DIR>(
            [&](int t) {
                const cpx a = sl[x.at(j + t * T)];
                if (t == 0) return a;
                return DIR < 0 ? cmulp(a, tw.w[t > 0 ? t - 1 : 0]) : cmulcp(a, tw.w[t > 0 ? t - 1 : 0]);
            },
            out);
    }
    x.after_load(xi0);
}

// number of shared-memory exchanges of one line transform
template <int N> constexpr int exchanges() { return 1; }

}  // namespace fast
}  // namespace psb
