// Atom -> slice binning and the structure-factor sum of the projected potential.
//
// Reference semantics (src/multislice/potentials.py):
//   :304-307  slice s holds atom a iff  lo[s] <= z_a < hi[s]  in float64, with
//             lo[s] = zs[s]-dz/2 (0 for s=0), hi[s] = zs[s]+dz/2 (zs[-1]+dz for the last slice);
//             the host evaluates lo[]/hi[] with those exact expressions, so neighbouring
//             intervals keep the reference's 1-ulp gaps/overlaps: an atom lands in 0, 1 or 2 slices.
//   :323-330  S_s[kx,ky] = sum_types f_Z[kx,ky] * sum_{a in s, type Z} e^{-2 pi i kx x_a} e^{-2 pi i ky y_a}
//
// Phases are reduced exactly: u = frac(x / L) is held as a 32-bit fixed-point fraction, so the
// integer product m*u wraps modulo one turn before it is converted to float.
#pragma once
#include "psb_common.cuh"

namespace psb {

struct BinParams {
    const double* pos;      // (F, A, 3) float64 positions
    const int* type_idx;    // (A) dense type index in [0, ntypes)
    int F, A, nz, ntypes;
    const double* lo;       // (nz)
    const double* hi;       // (nz)
    double inv_dz;          // 1 / (zs[1]-zs[0])
    double inv_lx, inv_ly;  // 1 / (nx*dx), 1 / (ny*dy)
    int* seg_of;            // (F, A, 2) segment ids (slice*ntypes + type) or -1
    int* offsets;           // (F, nseg+1): counts at [seg+1] after BinAssign, exclusive offsets after BinScan
    int* atom_list;         // (F, cap) atom indices grouped by segment, ascending inside a segment
    unsigned int* ux;       // (F, cap) frac(x/L) * 2^32
    unsigned int* uy;       // (F, cap)
    int cap;                // 2*A
};

// One thread per (atom, frame): exact float64 interval tests on the three candidate slices.
struct BinAssign {
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 1;
    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const BinParams& p) {
        const int a = cx.bx() * kThreads + cx.tid();
        const int f = cx.by();
        if (a >= p.A) return;
        const double z = p.pos[((long long)f * p.A + a) * 3 + 2];
        const int nseg = p.nz * p.ntypes;
        int found[2] = {-1, -1};
        int nf = 0;
        if (z == z) {   // NaN never matches, like the reference's comparisons
            double g = z * p.inv_dz + 0.5;
            int c = g < 0.0 ? 0 : (g > (double)(p.nz - 1) ? p.nz - 1 : (int)g);
            for (int s = c - 1; s <= c + 1; ++s) {
                if (s < 0 || s >= p.nz) continue;
                if (z >= p.lo[s] && z < p.hi[s] && nf < 2) found[nf++] = s;
            }
        }
        const int ty = p.type_idx[a];
        int* off = p.offsets + (long long)f * (nseg + 1);
        for (int i = 0; i < 2; ++i) {
            int seg = found[i] >= 0 ? found[i] * p.ntypes + ty : -1;
            p.seg_of[((long long)f * p.A + a) * 2 + i] = seg;
            if (seg >= 0) cx.atomic_add(&off[seg + 1], 1);
        }
    }
};

// One CTA per frame: counts -> exclusive offsets (in place; offsets[0] = 0).
struct BinScan {
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 1;
    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const BinParams& p) {
        const int nseg = p.nz * p.ntypes;
        int* off = p.offsets + (long long)cx.bx() * (nseg + 1);
        int* sm = reinterpret_cast<int*>(cx.smem());     // kThreads ints
        const int t = cx.tid();
        const int chunk = (nseg + kThreads - 1) / kThreads;
        const int i0 = 1 + t * chunk;
        const int i1 = (i0 + chunk < nseg + 1) ? i0 + chunk : nseg + 1;
        int s = 0;
        for (int i = i0; i < i1; ++i) s += off[i];
        sm[t] = s;
        cx.sync();
        int base = 0;
        for (int k = 0; k < t; ++k) base += sm[k];
        for (int i = i0; i < i1; ++i) {
            base += off[i];
            off[i] = base;
        }
        if (t == 0) off[0] = 0;
    }
};

// One CTA per (segment, frame): stable compaction of the atoms that belong to the segment, so the
// list order (ascending atom index) -- and with it every float32 sum downstream -- is deterministic.
struct BinCompact {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 1;
    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const BinParams& p) {
        const int seg = cx.bx(), f = cx.by();
        const int nseg = p.nz * p.ntypes;
        const int* off = p.offsets + (long long)f * (nseg + 1);
        const int begin = off[seg], count = off[seg + 1] - off[seg];
        if (count == 0) return;                       // block-uniform
        int* sm = reinterpret_cast<int*>(cx.smem());
        const int t = cx.tid();
        const int chunk = (p.A + kThreads - 1) / kThreads;
        const int a0 = t * chunk;
        const int a1 = a0 + chunk < p.A ? a0 + chunk : p.A;
        const int* so = p.seg_of + (long long)f * p.A * 2;
        int mine = 0;
        for (int a = a0; a < a1; ++a) mine += (so[2 * a] == seg) + (so[2 * a + 1] == seg);
        sm[t] = mine;
        cx.sync();
        int w = begin;
        for (int k = 0; k < t; ++k) w += sm[k];
        const double* pos = p.pos + (long long)f * p.A * 3;
        for (int a = a0; a < a1; ++a) {
            const int hits = (so[2 * a] == seg) + (so[2 * a + 1] == seg);
            for (int h = 0; h < hits; ++h) {
                const long long o = (long long)f * p.cap + w;
                p.atom_list[o] = a;
                double u = pos[3 * a] * p.inv_lx;
                double v = pos[3 * a + 1] * p.inv_ly;
                u -= floor(u);
                v -= floor(v);
                p.ux[o] = (unsigned int)((unsigned long long)(u * 4294967296.0 + 0.5) & 0xffffffffull);
                p.uy[o] = (unsigned int)((unsigned long long)(v * 4294967296.0 + 0.5) & 0xffffffffull);
                ++w;
            }
        }
    }
};

// exp(-2*pi*i*m*u) with u a 32-bit fraction: the integer product wraps modulo one turn exactly
PSB_D float2 unit_phase(int m, unsigned int u) {
    const int ph = (int)((unsigned int)m * u);                 // signed turn fraction * 2^32
    float s, c;
    sincospif((float)ph * 4.656612873077393e-10f, &s, &c);     // 2^-31: argument in [-1, 1)
    return make_float2(c, -s);
}

// ---- Hermitian, slice-paired structure factor ---------------------------------------------------
// V_s is real, so (i) only the half spectrum ky in [0, ny/2] is summed and the other half is written
// as its conjugate mirror, and (ii) two slices share one complex inverse FFT:
//     Z_m = S'_{2m} + i*S'_{2m+1}   ->   IFFT2(Z_m) = V_{2m} + i*V_{2m+1}.
// S' is the Hermitian part of the reference's spectrum: its Re(ifft2(S)) (potentials.py:336-337)
// discards the anti-Hermitian part, which is non-zero only on the self-conjugate Nyquist lines of an
// even-sized grid.  There e^{-2 pi i k x} is replaced by its real part (exactly the Hermitian part of
// the line); at the (Nyquist, Nyquist) corner the product's real part needs the extra -sin*sin term.
struct SfPairParams {
    const int* offsets;         // (F, nseg+1)
    const unsigned int* ux;     // (F, cap)
    const unsigned int* uy;
    int cap, nz, ntypes, nx, ny, npairs, pairs_per_block;
    const float* ff;            // (ntypes, nx, ny)
    float2* out;                // (F, npairs, nx, ny)
};

struct StructureFactorPaired {
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 2;
    static constexpr int TX = 64, TY = 32;   // kx x ky tile; thread: 4 kx (stride 16) x 2 ky (stride 16)
    static constexpr int CH = 32;
    static constexpr size_t kSmem = CH * (TX + TY) * sizeof(float2) + 2 * CH * sizeof(float);

    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const SfPairParams& p) {
        const int nyh = p.ny / 2 + 1;
        const int tiles_y = (nyh + TY - 1) / TY;
        const int kx0 = (cx.bx() / tiles_y) * TX, ky0 = (cx.bx() % tiles_y) * TY;
        const int f = cx.bz();
        const int tid = cx.tid(), tx = tid % 16, ty = tid / 16;
        float2* ex = reinterpret_cast<float2*>(cx.smem());     // [CH][TX]
        float2* ey = ex + CH * TX;                              // [CH][TY]
        float* snx = reinterpret_cast<float*>(ey + CH * TY);    // sin part of ex at the x Nyquist index
        float* sny = snx + CH;
        const int nseg = p.nz * p.ntypes;
        const int* off = p.offsets + (long long)f * (nseg + 1);
        const unsigned int* ux = p.ux + (long long)f * p.cap;
        const unsigned int* uy = p.uy + (long long)f * p.cap;
        const int hx = (p.nx + 1) / 2, hy = (p.ny + 1) / 2;
        const int nyq_x = (p.nx % 2 == 0) ? p.nx / 2 : -1;
        const int nyq_y = (p.ny % 2 == 0) ? p.ny / 2 : -1;
        const bool corner_tile = nyq_x >= kx0 && nyq_x < kx0 + TX && nyq_y >= ky0 && nyq_y < ky0 + TY;

        int kxs[4], kys[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) kxs[i] = kx0 + ty + 16 * i;
#pragma unroll
        for (int k = 0; k < 2; ++k) kys[k] = ky0 + tx + 16 * k;

        const int m0 = cx.by() * p.pairs_per_block;
        const int m1 = m0 + p.pairs_per_block < p.npairs ? m0 + p.pairs_per_block : p.npairs;
        for (int m = m0; m < m1; ++m) {
            float2 tot[2][4][2];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 2; ++k) tot[h][i][k] = make_float2(0.f, 0.f);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int s = 2 * m + h;
                if (s >= p.nz) continue;                          // block-uniform
                for (int t = 0; t < p.ntypes; ++t) {
                    const int b = off[s * p.ntypes + t], e = off[s * p.ntypes + t + 1];
                    if (b == e) continue;                          // block-uniform
                    float2 acc[4][2];
                    float corr = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 2; ++k) acc[i][k] = make_float2(0.f, 0.f);
                    for (int c0 = b; c0 < e; c0 += CH) {
                        const int nc = e - c0 < CH ? e - c0 : CH;
                        cx.sync();
                        for (int w = tid; w < nc * (TX + TY); w += kThreads) {
                            const int a = w / (TX + TY), r = w % (TX + TY);
                            if (r < TX) {
                                const int i = kx0 + r;
                                float2 z = unit_phase(i < hx ? i : i - p.nx, ux[c0 + a]);
                                if (i == nyq_x) { snx[a] = z.y; z.y = 0.f; }
                                ex[a * TX + r] = z;
                            } else {
                                const int i = ky0 + r - TX;
                                float2 z = unit_phase(i < hy ? i : i - p.ny, uy[c0 + a]);
                                if (i == nyq_y) { sny[a] = z.y; z.y = 0.f; }
                                ey[a * TY + r - TX] = z;
                            }
                        }
                        cx.sync();
                        for (int a = 0; a < nc; ++a) {
                            float2 xs[4], ys[2];
#pragma unroll
                            for (int i = 0; i < 4; ++i) xs[i] = ex[a * TX + ty + 16 * i];
#pragma unroll
                            for (int k = 0; k < 2; ++k) ys[k] = ey[a * TY + tx + 16 * k];
#pragma unroll
                            for (int i = 0; i < 4; ++i)
#pragma unroll
                                for (int k = 0; k < 2; ++k) {
                                    acc[i][k].x += xs[i].x * ys[k].x - xs[i].y * ys[k].y;
                                    acc[i][k].y += xs[i].x * ys[k].y + xs[i].y * ys[k].x;
                                }
                        }
                        if (corner_tile) {                         // block-uniform: Re(ex*ey) = cos*cos - sin*sin
                            for (int a = 0; a < nc; ++a) corr += snx[a] * sny[a];
                        }
                    }
                    const float* ff = p.ff + (long long)t * p.nx * p.ny;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            if (kxs[i] < p.nx && kys[k] < nyh) {
                                const float w = __ldg(&ff[(long long)kxs[i] * p.ny + kys[k]]);
                                float re = acc[i][k].x;
                                if (kxs[i] == nyq_x && kys[k] == nyq_y) re -= corr;
                                tot[h][i][k].x += re * w;
                                tot[h][i][k].y += acc[i][k].y * w;
                            }
                        }
                }
            }
            // Z[k] = A + iB,  Z[-k] = conj(A) + i*conj(B)
            float2* out = p.out + ((long long)f * p.npairs + m) * p.nx * p.ny;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int kx = kxs[i], ky = kys[k];
                    if (kx < p.nx && ky < nyh) {
                        const float2 A = tot[0][i][k], B = tot[1][i][k];
                        out[(long long)kx * p.ny + ky] = make_float2(A.x - B.y, A.y + B.x);
                        const int my = ky == 0 ? 0 : p.ny - ky;
                        if (my >= nyh) {
                            const int mx = kx == 0 ? 0 : p.nx - kx;
                            out[(long long)mx * p.ny + my] = make_float2(A.x + B.y, B.x - A.y);
                        }
                    }
                }
        }
    }
};

// t = exp(i*sigma*V) from a real potential (for Propagate() on a user-supplied Potential object)
struct TransmitParams {
    const float* V;
    float2* t;
    long long n;
    float sigma;
};
struct Transmit {
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 1;
    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const TransmitParams& p) {
        for (long long i = (long long)cx.bx() * kThreads + cx.tid(); i < p.n; i += (long long)cx.gx() * kThreads) {
            float s, c;
            sincosf(p.sigma * p.V[i], &s, &c);
            p.t[i] = make_float2(c, s);
        }
    }
};

}  // namespace psb
