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
    int* cursor;            // (F, nseg) fill counters of BinScatter (zeroed by the host)
    int* unsorted;          // (F, cap) atom indices grouped by segment, arrival order
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

// Compaction in two steps.  BinScatter drops every (atom, hit) into its segment in arrival order (atomic
// cursor per segment); BinOrder, one warp per (segment, frame), ranks the handful of atoms of the segment by
// atom index and writes them out sorted together with their fixed-point x/y fractions, so the list order
// (ascending atom index) -- and with it every float32 sum downstream -- is deterministic.
struct BinScatter {
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 1;
    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const BinParams& p) {
        const int a = cx.bx() * kThreads + cx.tid();
        const int f = cx.by();
        if (a >= p.A) return;
        const int nseg = p.nz * p.ntypes;
        const int* off = p.offsets + (long long)f * (nseg + 1);
        int* cur = p.cursor + (long long)f * nseg;
        int* tmp = p.unsorted + (long long)f * p.cap;
        for (int i = 0; i < 2; ++i) {
            const int seg = p.seg_of[((long long)f * p.A + a) * 2 + i];
            if (seg >= 0) tmp[off[seg] + cx.atomic_add(&cur[seg], 1)] = a;
        }
    }
};

// segments up to this many atoms are ranked by one warp (BinOrder); larger ones by a whole CTA through shared memory
// (BinOrderLarge): the all-pairs ranking is O(n^2), fine for the tens of atoms of a bulk-crystal slice, a cliff for the
// 10^3..10^5 atoms a 2-D material or a thick slice puts into one (slice, type) segment
constexpr int kBinOrderWarpMax = 64;

PSB_D void bin_emit(const BinParams& p, int f, long long o, int a) {
    const double* pos = p.pos + (long long)f * p.A * 3;
    p.atom_list[o] = a;
    double u = pos[3 * a] * p.inv_lx;
    double v = pos[3 * a + 1] * p.inv_ly;
    u -= floor(u);
    v -= floor(v);
    p.ux[o] = (unsigned int)((unsigned long long)(u * 4294967296.0 + 0.5) & 0xffffffffull);
    p.uy[o] = (unsigned int)((unsigned long long)(v * 4294967296.0 + 0.5) & 0xffffffffull);
}

struct BinOrder {
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 1;
    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const BinParams& p) {
        const int nseg = p.nz * p.ntypes;
        const int seg = cx.bx() * (kThreads / 32) + cx.tid() / 32, f = cx.by();
        if (seg >= nseg) return;
        const int lane = cx.tid() % 32;
        const int* off = p.offsets + (long long)f * (nseg + 1);
        const int begin = off[seg], n = off[seg + 1] - off[seg];
        if (n > kBinOrderWarpMax) return;                         // BinOrderLarge's
        const int* tmp = p.unsorted + (long long)f * p.cap + begin;
        for (int i = lane; i < n; i += 32) {
            const int a = tmp[i];
            int rank = 0;
            for (int k = 0; k < n; ++k) rank += tmp[k] < a;       // an atom occurs at most once per segment
            bin_emit(p, f, (long long)f * p.cap + begin + rank, a);
        }
    }
};

// One CTA per (segment, frame) with more than kBinOrderWarpMax atoms: the ids are staged in shared memory in chunks and
// every thread ranks kPer atoms at a time against a chunk (one broadcast read per kPer compares).
struct BinOrderLarge {
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 1;
    static constexpr int kChunk = 2048;
    static constexpr int kPer = 4;
    static constexpr size_t kSmem = kChunk * sizeof(int);
    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const BinParams& p) {
        const int nseg = p.nz * p.ntypes;
        const int seg = cx.bx(), f = cx.by();
        const int* off = p.offsets + (long long)f * (nseg + 1);
        const int begin = off[seg], n = off[seg + 1] - off[seg];
        if (n <= kBinOrderWarpMax) return;
        const int* tmp = p.unsorted + (long long)f * p.cap + begin;
        int* sm = reinterpret_cast<int*>(cx.smem());
        const int t = cx.tid();
        for (int a0 = 0; a0 < n; a0 += kThreads * kPer) {
            int a[kPer], rank[kPer];
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int idx = a0 + i * kThreads + t;
                a[i] = idx < n ? tmp[idx] : 0x7fffffff;
                rank[i] = 0;
            }
            for (int c0 = 0; c0 < n; c0 += kChunk) {
                const int m = n - c0 < kChunk ? n - c0 : kChunk;
                cx.sync();
                for (int k = t; k < m; k += kThreads) sm[k] = tmp[c0 + k];
                cx.sync();
                for (int k = 0; k < m; ++k) {
                    const int v = sm[k];
#pragma unroll
                    for (int i = 0; i < kPer; ++i) rank[i] += v < a[i];
                }
            }
#pragma unroll
            for (int i = 0; i < kPer; ++i)
                if (a0 + i * kThreads + t < n) bin_emit(p, f, (long long)f * p.cap + begin + rank[i], a[i]);
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

// ---- quarter-spectrum, slice-paired structure factor ----------------------------------------------
// V_s is real and e^{-2 pi i k x} = c - i*s with c even / s odd in k, so with the four real sums
//     CC = sum cx*cy,  SS = sum sx*sy,  CS = sum cx*sy,  SC = sum sx*cy      (4 FMA per atom and slot)
// over the atoms of a (slice, type), one slot (kx, ky) with kx, ky >= 0 yields all four spectrum entries
//     S[+kx,+ky] = (CC-SS) - i(CS+SC)        S[-kx,+ky] = (CC+SS) - i(CS-SC)
//     S[+kx,-ky] = conj S[-kx,+ky]           S[-kx,-ky] = conj S[+kx,+ky]
// i.e. one real FMA per complex spectrum entry instead of four.  Two slices then share one complex
// inverse FFT: Z_m = S'_{2m} + i*S'_{2m+1}  ->  IFFT2(Z_m) = V_{2m} + i*V_{2m+1}.
//
// S' is the Hermitian part of the reference's spectrum: its Re(ifft2(S)) (potentials.py:336-337) drops the
// anti-Hermitian part, which is non-zero only on the self-conjugate Nyquist lines of an even-sized axis,
// where the Hermitian part of e^{-2 pi i k x} is its cosine.  The zero-frequency slot has s = 0, so the
// Nyquist cosine rides in that unused component: slot 0 stores (1, cos(pi*n*u)) and its SC / SS sums are
// the Nyquist line's CC / CS.  Only the (Nyquist, Nyquist) corner needs one extra term, -sum sin*sin.
struct SfPairParams {
    const int* offsets;         // (F, nseg+1)
    const unsigned int* ux;     // (F, cap)
    const unsigned int* uy;
    int cap, nz, ntypes, nx, ny;
    int pair_begin, pair_count; // this launch covers slice pairs [pair_begin, pair_begin + pair_count) of every frame
    int pairs_per_block;
    const float* ff;            // (ntypes, nx, ny)
    float2* out;                // (F, pair_count, nx, ny)
};

struct StructureFactorPaired {
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 2;
    static constexpr int TX = 64, TY = 32;   // slots per tile; thread: 4 kx slots (stride 16) x 2 ky slots (stride 16)
    static constexpr int CH = 32;            // atoms staged per chunk
    static constexpr size_t kSmem = CH * (TX + TY) * sizeof(float2) + 2 * CH * sizeof(float) + 32 * kThreads * sizeof(float);

    // number of non-negative-frequency slots of an axis (the Nyquist line of an even axis rides in slot 0)
    static PSB_HD int slots(int n) { return (n % 2 == 0) ? n / 2 : (n + 1) / 2; }

    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const SfPairParams& p) {
        const int nsx = slots(p.nx), nsy = slots(p.ny);
        const int tiles_y = (nsy + TY - 1) / TY;
        const int kx0 = (cx.bx() / tiles_y) * TX, ky0 = (cx.bx() % tiles_y) * TY;
        const int f = cx.bz();
        const int tid = cx.tid(), tx = tid % 16, ty = tid / 16;
        float2* ex = reinterpret_cast<float2*>(cx.smem());     // [CH][TX]  (cx, sx)
        float2* ey = ex + CH * TX;                              // [CH][TY]  (cy, sy)
        float* snx = reinterpret_cast<float*>(ey + CH * TY);    // sin(pi*nx*u) per staged atom (corner term)
        float* sny = snx + CH;
        float* stash = sny + CH;                                // [32][kThreads] totals of the pair's first slice
        const int nseg = p.nz * p.ntypes;
        const int* off = p.offsets + (long long)f * (nseg + 1);
        const unsigned int* ux = p.ux + (long long)f * p.cap;
        const unsigned int* uy = p.uy + (long long)f * p.cap;
        const bool nq_x = p.nx % 2 == 0, nq_y = p.ny % 2 == 0;
        const bool corner_tile = kx0 == 0 && ky0 == 0 && nq_x && nq_y;

        int gx[4], gy[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) gx[i] = kx0 + ty + 16 * i;
#pragma unroll
        for (int k = 0; k < 2; ++k) gy[k] = ky0 + tx + 16 * k;

        const int l0 = cx.by() * p.pairs_per_block;
        const int l1 = l0 + p.pairs_per_block < p.pair_count ? l0 + p.pairs_per_block : p.pair_count;
        for (int ml = l0; ml < l1; ++ml) {
            const int m = p.pair_begin + ml;
            float tot[4][2][4];     // [i][k][cc, ss, cs, sc], already multiplied by the form factor
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 2; ++k)
#pragma unroll
                        for (int q = 0; q < 4; ++q) tot[i][k][q] = 0.f;
                const int s = 2 * m + h;
                if (s < p.nz) {                                   // block-uniform
                    for (int t = 0; t < p.ntypes; ++t) {
                        const int b = off[s * p.ntypes + t], e = off[s * p.ntypes + t + 1];
                        if (b == e) continue;                      // block-uniform
                        float acc[4][2][4];
                        float corr = 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int k = 0; k < 2; ++k)
#pragma unroll
                                for (int q = 0; q < 4; ++q) acc[i][k][q] = 0.f;
                        for (int c0 = b; c0 < e; c0 += CH) {
                            const int nc = e - c0 < CH ? e - c0 : CH;
                            cx.sync();
                            for (int w = tid; w < nc * (TX + TY); w += kThreads) {
                                const int a = w / (TX + TY), r = w % (TX + TY);
                                if (r < TX) {
                                    const int g = kx0 + r;
                                    float2 z = unit_phase(g, ux[c0 + a]);          // (c, -s)
                                    z.y = -z.y;
                                    if (g == 0 && nq_x) {
                                        const float2 n = unit_phase(p.nx / 2, ux[c0 + a]);
                                        z.y = n.x;
                                        snx[a] = n.y;
                                    }
                                    ex[a * TX + r] = z;
                                } else {
                                    const int g = ky0 + r - TX;
                                    float2 z = unit_phase(g, uy[c0 + a]);
                                    z.y = -z.y;
                                    if (g == 0 && nq_y) {
                                        const float2 n = unit_phase(p.ny / 2, uy[c0 + a]);
                                        z.y = n.x;
                                        sny[a] = n.y;
                                    }
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
                                        acc[i][k][0] += xs[i].x * ys[k].x;
                                        acc[i][k][1] += xs[i].y * ys[k].y;
                                        acc[i][k][2] += xs[i].x * ys[k].y;
                                        acc[i][k][3] += xs[i].y * ys[k].x;
                                    }
                            }
                            if (corner_tile)                        // block-uniform
                                for (int a = 0; a < nc; ++a) corr += snx[a] * sny[a];
                        }
                        const float* ff = p.ff + (long long)t * p.nx * p.ny;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                if (gx[i] < nsx && gy[k] < nsy) {
                                    // row / column of the form-factor table for the cosine and the "sine" component
                                    const int fx0 = gx[i], fx1 = (gx[i] == 0 && nq_x) ? p.nx / 2 : gx[i];
                                    const int fy0 = gy[k], fy1 = (gy[k] == 0 && nq_y) ? p.ny / 2 : gy[k];
                                    float ss = acc[i][k][1];
                                    if (corner_tile && gx[i] == 0 && gy[k] == 0) ss -= corr;
                                    tot[i][k][0] += acc[i][k][0] * __ldg(&ff[(long long)fx0 * p.ny + fy0]);
                                    tot[i][k][1] += ss * __ldg(&ff[(long long)fx1 * p.ny + fy1]);
                                    tot[i][k][2] += acc[i][k][2] * __ldg(&ff[(long long)fx0 * p.ny + fy1]);
                                    tot[i][k][3] += acc[i][k][3] * __ldg(&ff[(long long)fx1 * p.ny + fy0]);
                                }
                            }
                    }
                }
                if (h == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 2; ++k)
#pragma unroll
                            for (int q = 0; q < 4; ++q) stash[((i * 2 + k) * 4 + q) * kThreads + tid] = tot[i][k][q];
                }
            }
            // tot = second slice (B), stash = first slice (A):  Z = S'_A + i*S'_B at up to four mirror positions
            float2* out = p.out + ((long long)f * p.pair_count + ml) * p.nx * p.ny;
            auto emit = [&](int kx, int ky, float ar, float ai, float br, float bi) {
                out[(long long)kx * p.ny + ky] = make_float2(ar - bi, ai + br);
            };
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int x = gx[i], y = gy[k];
                    if (x >= nsx || y >= nsy) continue;
                    float A[4], B[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        A[q] = stash[((i * 2 + k) * 4 + q) * kThreads + tid];
                        B[q] = tot[i][k][q];
                    }
                    const float acc_ = A[0], ass = A[1], acs = A[2], asc = A[3];
                    const float bcc = B[0], bss = B[1], bcs = B[2], bsc = B[3];
                    if (x > 0 && y > 0) {
                        emit(x, y, acc_ - ass, -(acs + asc), bcc - bss, -(bcs + bsc));
                        emit(p.nx - x, y, acc_ + ass, -(acs - asc), bcc + bss, -(bcs - bsc));
                        emit(x, p.ny - y, acc_ + ass, acs - asc, bcc + bss, bcs - bsc);
                        emit(p.nx - x, p.ny - y, acc_ - ass, acs + asc, bcc - bss, bcs + bsc);
                    } else if (x == 0 && y > 0) {
                        emit(0, y, acc_, -acs, bcc, -bcs);
                        emit(0, p.ny - y, acc_, acs, bcc, bcs);
                        if (nq_x) {
                            emit(p.nx / 2, y, asc, -ass, bsc, -bss);
                            emit(p.nx / 2, p.ny - y, asc, ass, bsc, bss);
                        }
                    } else if (x > 0 && y == 0) {
                        emit(x, 0, acc_, -asc, bcc, -bsc);
                        emit(p.nx - x, 0, acc_, asc, bcc, bsc);
                        if (nq_y) {
                            emit(x, p.ny / 2, acs, -ass, bcs, -bss);
                            emit(p.nx - x, p.ny / 2, acs, ass, bcs, bss);
                        }
                    } else {
                        emit(0, 0, acc_, 0.f, bcc, 0.f);
                        if (nq_y) emit(0, p.ny / 2, acs, 0.f, bcs, 0.f);
                        if (nq_x) emit(p.nx / 2, 0, asc, 0.f, bsc, 0.f);
                        if (nq_x && nq_y) emit(p.nx / 2, p.ny / 2, ass, 0.f, bss, 0.f);
                    }
                }
        }
    }
};

// t = exp(i*sigma*V) from a real potential (for Propagate() on a user-supplied Potential object)
// n elements in blocks of `block` contiguous ones: block b of V starts at b * v_block_stride, of t at b * block
// (block = n, one block: a plain array; block = one image, v_block_stride = one frame's stack: slice 0 of every frame)
struct TransmitParams {
    const float* V;
    float2* t;
    long long n;
    float sigma;
    long long block, v_block_stride;
};
struct Transmit {
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 1;
    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const TransmitParams& p) {
        for (long long i = (long long)cx.bx() * kThreads + cx.tid(); i < p.n; i += (long long)cx.gx() * kThreads) {
            const long long b = i / p.block;
            float s, c;
            sincosf(p.sigma * p.V[b * p.v_block_stride + (i - b * p.block)], &s, &c);
            p.t[i] = make_float2(c, s);
        }
    }
};

}  // namespace psb
