// EXPERIMENT, not part of libpsb: sf_fast.cu with sf_tiles_persistent_kernel (PSB_SF_PERSISTENT) -- two resident CTAs per SM walk a list of
// (pair image, tile) items with the offsets of all items staged once and the table ring running across item boundaries.
// Bit-identical, but slower than one CTA per item: C2 2.366 against 2.160 ms per 16 frames, C4 direct sum 22.6 against 17.7 ms
// per 8 frames (profiles/r2ah_sf_persistent_tiles.txt): the tiles kernel is bound by its instruction count (2000 per warp and
// item, 31 % of them the FFMA2 loop), not by the latency chain of a CTA's start-up.
// Structure-factor sum of the projected potential, pipelined version (reference: src/multislice/potentials.py
// :319-330, the einsum over exp(-2 pi i kx x) exp(-2 pi i ky y) of the atoms of a slice, times the form factor).
//
// Same arithmetic as StructureFactorPaired (potential_kernels.cuh: quarter spectrum, four real sums per slot,
// two slices packed per complex image, Nyquist lines riding in slot 0), restructured so the hot loop is nothing
// but shared-memory loads and packed FMAs:
//
//   K1  PhaseTables   once per chunk: (cos, sin)(2 pi g u) of every atom entry of the chunk for every
//                     non-negative frequency slot g of both axes, with the exact 32-bit fixed-point phase
//                     reduction, laid out [tile][entry][slot-in-tile] so that the rows a tile needs for one
//                     (slice, type) segment are ONE contiguous block;
//   K2  SfTiles       one CTA per (64 x 32 slot tile, slice pair, frame): the blocks of up to 32 atoms stream
//                     global -> shared with cp.async.bulk through a 2-stage mbarrier ring (deeper rings measured: no gain) while the previous
//                     block is accumulated with FFMA2 (a thread owns 4 x 2 slots x 4 sums = 16 packed
//                     accumulators); form factor, pairing and the mirror expansion as before.
//
// The tables stay L2-resident between K1 and K2 (a few MB per chunk).
#include "fast_fft.cuh"
#include "fast_path.h"
#include "graph_cache.h"
#include "pdl.cuh"
#include "potential_kernels.cuh"
#include "psb_rt.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

namespace psb {

namespace {

using fast::cpx;

constexpr int TX = 64, TY = 32, CH = 32;      // slots per tile along kx / ky, atoms per staged block

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_%=;\n"
        " bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct SfFastParams {
    const int* offsets;         // (nf, nseg+1), first frame of the chunk
    const unsigned int* ux;     // (nf, cap)
    const unsigned int* uy;
    int cap, nz, ntypes, nx, ny;
    int pair_begin, pair_count;
    int tiles_x, tiles_y;
    float2* tabx;               // [nf][tiles_x][cap][TX]   (cos, sin)(2 pi g u); slot 0 of an even axis: (1, cos(pi n u))
    float2* taby;               // [nf][tiles_y][cap][TY]
    float* snx;                 // [nf][cap]  Nyquist sine (corner term), even axes only
    float* sny;
    const float4* ff4;          // (ntypes, nsx, nsy): form factor at the four table positions a slot's sums need
    float2* out;                // (nf, pair_count, nx, ny)
};

// (cos, sin) of 2*pi*g*u for one axis entry, laid out as StructureFactorPaired stages it.  u is a 32-bit turn
// fraction, so g*u wraps exactly and the angle handed to the SFU (sin.approx / cos.approx) lies in [-pi, pi), where
// their absolute error is <= 2^-21.2; the potential's error against the oracle goes from 1.8e-7 to 3.2e-7 rel-L2
// (budget 1e-5).  libdevice's sincospif made this kernel ALU-bound (ncu r1h: 60 % ALU pipe, 6.7 us per chunk).
__device__ __forceinline__ float2 sfu_phase(int g, unsigned int u) {
    const int ph = (int)((unsigned int)g * u);                 // signed turn fraction * 2^32
    float s, c;
    __sincosf((float)ph * 1.4629180792671596e-9f, &s, &c);     // pi * 2^-31
    return make_float2(c, s);
}
__device__ __forceinline__ float2 slot_phase(int g, unsigned int u, int n, bool nyq, float* sn_out) {
    float2 z = sfu_phase(g, u);
    if (g == 0 && nyq) {
        const float2 q = sfu_phase(n / 2, u);
        z.y = q.x;
        *sn_out = -q.y;
    }
    return z;
}

// K0 (once per psb_build_transmission call): gather the form factor for the four real sums of every slot
//   .x -> cc at (kx, ky)   .y -> ss at (kx', ky')   .z -> cs at (kx, ky')   .w -> sc at (kx', ky)
// where kx' = kx except for slot 0 of an even axis, whose sine component carries the Nyquist line (kx' = n/2)
__global__ void __launch_bounds__(256) ff4_kernel(const float* ff, float4* ff4, int ntypes, int nx, int ny) {
    const int nsx = StructureFactorPaired::slots(nx), nsy = StructureFactorPaired::slots(ny);
    const long long n = (long long)ntypes * nsx * nsy;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int gy = (int)(i % nsy), gx = (int)((i / nsy) % nsx), t = (int)(i / ((long long)nsx * nsy));
        const int fx1 = (gx == 0 && nx % 2 == 0) ? nx / 2 : gx, fy1 = (gy == 0 && ny % 2 == 0) ? ny / 2 : gy;
        const float* f = ff + (long long)t * nx * ny;
        ff4[i] = make_float4(f[(long long)gx * ny + gy], f[(long long)fx1 * ny + fy1], f[(long long)gx * ny + fy1],
                             f[(long long)fx1 * ny + gy]);
    }
}

// K1: one warp per atom entry of the chunk
__global__ void __launch_bounds__(256) phase_tables_kernel(const SfFastParams p) {
    pdl_trigger();
    pdl_wait();          // the previous chunk's kernels may still be reading the tables
    const int fl = blockIdx.y;
    const int nseg = p.nz * p.ntypes;
    const int* off = p.offsets + (long long)fl * (nseg + 1);
    const int s_begin = 2 * p.pair_begin;
    int s_end = 2 * (p.pair_begin + p.pair_count);
    if (s_end > p.nz) s_end = p.nz;
    const int e_begin = off[s_begin * p.ntypes], e_end = off[s_end * p.ntypes];
    const int e = e_begin + blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= e_end) return;
    const int lane = threadIdx.x & 31;
    const unsigned int ux = p.ux[(long long)fl * p.cap + e], uy = p.uy[(long long)fl * p.cap + e];
    const int nsx = StructureFactorPaired::slots(p.nx), nsy = StructureFactorPaired::slots(p.ny);
    const bool nqx = p.nx % 2 == 0, nqy = p.ny % 2 == 0;
    float sn = 0.f;
    for (int g = lane; g < p.tiles_x * TX; g += 32) {
        float2 z = g < nsx ? slot_phase(g, ux, p.nx, nqx, &sn) : make_float2(0.f, 0.f);
        p.tabx[(((long long)fl * p.tiles_x + g / TX) * p.cap + e) * TX + g % TX] = z;
    }
    if (lane == 0 && nqx) p.snx[(long long)fl * p.cap + e] = sn;
    for (int g = lane; g < p.tiles_y * TY; g += 32) {
        float2 z = g < nsy ? slot_phase(g, uy, p.ny, nqy, &sn) : make_float2(0.f, 0.f);
        p.taby[(((long long)fl * p.tiles_y + g / TY) * p.cap + e) * TY + g % TY] = z;
    }
    if (lane == 0 && nqy) p.sny[(long long)fl * p.cap + e] = sn;
}

// iterator over the staged blocks of one slice pair, in the order (slice of the pair, type, block)
struct BlockIter {
    int h, t, c0, e;            // current slice-of-pair, type, block start, segment end
    bool valid;
};

__device__ __forceinline__ void iter_seek(BlockIter& it, const int* off, int m, int nz, int ntypes) {
    // position on the first non-empty segment at or after (it.h, it.t); c0 = its start
    while (it.h < 2) {
        const int s = 2 * m + it.h;
        if (s < nz) {
            while (it.t < ntypes) {
                const int b = off[s * ntypes + it.t], e = off[s * ntypes + it.t + 1];
                if (b < e) {
                    it.c0 = b;
                    it.e = e;
                    it.valid = true;
                    return;
                }
                ++it.t;
            }
        }
        ++it.h;
        it.t = 0;
    }
    it.valid = false;
}
__device__ __forceinline__ void iter_next(BlockIter& it, const int* off, int m, int nz, int ntypes) {
    it.c0 += CH;
    if (it.c0 < it.e) return;
    ++it.t;
    iter_seek(it, off, m, nz, ntypes);
}

constexpr int kStageElems = CH * (TX + TY);
constexpr int kMaxStagedTypes = 64;
#ifndef PSB_SF_STAGES
#define PSB_SF_STAGES 2
#endif
constexpr int kSfStages = PSB_SF_STAGES;      // blocks of atoms in flight per CTA (ring depth)
constexpr size_t kSfSmem = kSfStages * (size_t)kStageElems * sizeof(float2) + kSfStages * sizeof(uint64_t) + 32 * 256 * sizeof(float);

// K2
__global__ void __launch_bounds__(256, 2) sf_tiles_kernel(const SfFastParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* stage = reinterpret_cast<cpx*>(smem_raw);                                  // [kSfStages][CH*TX + CH*TY]
    uint64_t* full = reinterpret_cast<uint64_t*>(stage + kSfStages * kStageElems);  // [kSfStages]
    // totals of the pair's first slice wait here while the second one accumulates: as 32 more live registers they
    // pushed the kernel to the 128-register cap and ptxas stopped hoisting the loop's loads over its FMAs (ncu r1i,
    // C4 geometry: short-scoreboard stalls on every FFMA2 of that copy of the loop, 2.3x the time per atom)
    float* stash = reinterpret_cast<float*>(full + kSfStages);                      // [32][256]

    const int nsx = StructureFactorPaired::slots(p.nx), nsy = StructureFactorPaired::slots(p.ny);
    const int tile_x = blockIdx.x / p.tiles_y, tile_y = blockIdx.x % p.tiles_y;
    const int kx0 = tile_x * TX, ky0 = tile_y * TY;
    const int ml = blockIdx.y, m = p.pair_begin + ml, fl = blockIdx.z;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int nseg = p.nz * p.ntypes;
    const int* goff = p.offsets + (long long)fl * (nseg + 1);
    const bool nq_x = p.nx % 2 == 0, nq_y = p.ny % 2 == 0;
    const bool corner_tile = kx0 == 0 && ky0 == 0 && nq_x && nq_y;
    const float2* tabx = p.tabx + ((long long)fl * p.tiles_x + tile_x) * p.cap * TX;
    const float2* taby = p.taby + ((long long)fl * p.tiles_y + tile_y) * p.cap * TY;
    const float* snx = p.snx + (long long)fl * p.cap;
    const float* sny = p.sny + (long long)fl * p.cap;

    pdl_trigger();
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kSfStages; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();          // offsets / tables come from the previous kernels of the chain
    // the pair's segment offsets, staged once: the producer's block iterator and every thread's segment loop read
    // them many times, and as global loads each of those reads was a dependent L2 round trip (ncu r1h: ~10 % of the
    // kernel's stall samples)
    __shared__ int soff[2 * kMaxStagedTypes + 1];          // launch_sf_fast rejects more types
    const int obase = 2 * m * p.ntypes;
    if (tid <= 2 * p.ntypes) soff[tid] = goff[obase + tid < nseg ? obase + tid : nseg];
    const int* off = soff - obase;
    __syncthreads();

    int gx[4], gy[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) gx[i] = kx0 + ty + 16 * i;
#pragma unroll
    for (int k = 0; k < 2; ++k) gy[k] = ky0 + tx + 16 * k;

    // producer side: thread 0 keeps kSfStages blocks in flight
    BlockIter ahead{0, 0, 0, 0, false};
    iter_seek(ahead, off, m, p.nz, p.ntypes);
    int issued = 0;
    auto issue = [&]() {
        if (!ahead.valid) return;
        if (tid == 0) {
            const int nc = ahead.e - ahead.c0 < CH ? ahead.e - ahead.c0 : CH;
            cpx* dst = stage + (issued % kSfStages) * kStageElems;
            uint64_t* bar = &full[issued % kSfStages];
            mbar_expect_tx(bar, (uint32_t)(nc * (TX + TY) * sizeof(float2)));
            bulk_g2s(dst, tabx + (long long)ahead.c0 * TX, (uint32_t)(nc * TX * sizeof(float2)), bar);
            bulk_g2s(dst + CH * TX, taby + (long long)ahead.c0 * TY, (uint32_t)(nc * TY * sizeof(float2)), bar);
        }
        ++issued;
        iter_next(ahead, off, m, p.nz, p.ntypes);
    };
#pragma unroll 1
    for (int i = 0; i < kSfStages; ++i) issue();

    int used = 0;
    float tot[4][2][4];     // [i][k][cc, ss, cs, sc], multiplied by the form factor
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int q = 0; q < 4; ++q) tot[i][k][q] = 0.f;
        const int s = 2 * m + h;
        if (s < p.nz) {
            for (int t = 0; t < p.ntypes; ++t) {
                const int b = off[s * p.ntypes + t], e = off[s * p.ntypes + t + 1];
                if (b == e) continue;
                cpx P[4][2], Q[4][2];      // P = (cc, cs), Q = (sc, ss)
                float corr = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        P[i][k] = fast::c_make(0.f, 0.f);
                        Q[i][k] = fast::c_make(0.f, 0.f);
                    }
                for (int c0 = b; c0 < e; c0 += CH) {
                    const int nc = e - c0 < CH ? e - c0 : CH;
                    const cpx* ex = stage + (used % kSfStages) * kStageElems;
                    const cpx* ey = ex + CH * TX;
                    mbar_wait(&full[used % kSfStages], (uint32_t)((used / kSfStages) & 1));
#pragma unroll 4
                    for (int a = 0; a < nc; ++a) {
                        cpx ys[2];
#pragma unroll
                        for (int k = 0; k < 2; ++k) ys[k] = ey[a * TY + tx + 16 * k];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const cpx x = ex[a * TX + ty + 16 * i];
                            const cpx xc = fast::c_make(fast::c_re(x), fast::c_re(x));
                            const cpx xs = fast::c_make(fast::c_im(x), fast::c_im(x));
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                P[i][k] = fast::fma2(xc, ys[k], P[i][k]);
                                Q[i][k] = fast::fma2(xs, ys[k], Q[i][k]);
                            }
                        }
                    }
                    if (corner_tile)
                        for (int a = 0; a < nc; ++a) corr += __ldg(&snx[c0 + a]) * __ldg(&sny[c0 + a]);
                    __syncthreads();             // every thread is done with this stage
                    ++used;
                    issue();                     // refill it with the block after the one already in flight
                }
                const float4* ff4 = p.ff4 + (long long)t * nsx * nsy;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        if (gx[i] < nsx && gy[k] < nsy) {
                            const float4 f = __ldg(&ff4[gx[i] * nsy + gy[k]]);
                            const float cc = fast::c_re(P[i][k]), cs = fast::c_im(P[i][k]);
                            const float sc = fast::c_re(Q[i][k]);
                            float ss = fast::c_im(Q[i][k]);
                            if (corner_tile && gx[i] == 0 && gy[k] == 0) ss -= corr;
                            tot[i][k][0] += cc * f.x;
                            tot[i][k][1] += ss * f.y;
                            tot[i][k][2] += cs * f.z;
                            tot[i][k][3] += sc * f.w;
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
                    for (int q = 0; q < 4; ++q) stash[((i * 2 + k) * 4 + q) * 256 + tid] = tot[i][k][q];
        }
    }
    // tot = second slice (B), first = first slice (A):  Z = S'_A + i*S'_B at up to four mirror positions
    float2* out = p.out + ((long long)fl * p.pair_count + ml) * p.nx * p.ny;
    auto emit = [&](int kx, int ky, float ar, float ai, float br, float bi) {
        out[kx * p.ny + ky] = make_float2(ar - bi, ai + br);      // nx*ny < 2^31
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
                A[q] = stash[((i * 2 + k) * 4 + q) * 256 + tid];
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

// K2, persistent form.  One launch of K2 above is pair_count * tiles CTAs that each live ~9 us for ~2 us of work at C2's ~39
// atoms per pair: a chain of dependent round trips (segment offsets -> first table block -> form factors of slice A -> form
// factors of slice B), paid by every CTA (ncu r2ag: 15 / 23 / 27 / 38 us for 1 / 2 / 3 / 4 rounds of CTAs).  Here 2 CTAs per SM
// stay resident and walk a list of work items (item = pair image * tiles + tile, stride = grid): the offsets of all of a
// CTA's items are staged once, the ring of table blocks runs on across item boundaries (the producer is always two blocks
// ahead, whatever pair they belong to), and with the grid a multiple of the tile count a CTA keeps its tile, so the form
// factors of its slots come from L1 after the first item.  Same arithmetic, same order of summation: bit-identical.
#ifndef PSB_SF_PERSISTENT
#define PSB_SF_PERSISTENT 1
#endif
constexpr int kSfMaxItems = 16;               // work items per CTA and launch (the host splits larger chunks)

struct ItemIter {
    int k;                      // index into the CTA's item list
    int h, t, c0, e;            // slice of the pair, type, block start, segment end
    bool valid;
};

__global__ void __launch_bounds__(256, 2) sf_tiles_persistent_kernel(const SfFastParams p, int n_items) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* stage = reinterpret_cast<cpx*>(smem_raw);                                  // [kSfStages][CH*TX + CH*TY]
    uint64_t* full = reinterpret_cast<uint64_t*>(stage + kSfStages * kStageElems);  // [kSfStages]
    float* stash = reinterpret_cast<float*>(full + kSfStages);                      // [32][256]
    __shared__ int soff[kSfMaxItems][2 * kMaxStagedTypes + 1];                      // segment offsets of the CTA's items

    const int nsx = StructureFactorPaired::slots(p.nx), nsy = StructureFactorPaired::slots(p.ny);
    const int tiles = p.tiles_x * p.tiles_y;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int nseg = p.nz * p.ntypes;
    const bool nq_x = p.nx % 2 == 0, nq_y = p.ny % 2 == 0;
    const int G = gridDim.x;
    const int first = blockIdx.x;
    const int mine = first < n_items ? (n_items - first + G - 1) / G : 0;           // <= kSfMaxItems (host)

    pdl_trigger();
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kSfStages; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_wait();          // offsets / tables come from the previous kernels of the chain
    auto item_pair = [&](int k, int* fl, int* m, int* tile) {
        const int item = first + k * G;
        const int q = item / tiles;
        *tile = item - q * tiles;
        *fl = q / p.pair_count;
        *m = p.pair_begin + (q - *fl * p.pair_count);
    };
    for (int i = tid; i < mine * (2 * p.ntypes + 1); i += 256) {
        const int k = i / (2 * p.ntypes + 1), r = i - k * (2 * p.ntypes + 1);
        int fl, m, tile;
        item_pair(k, &fl, &m, &tile);
        const int* goff = p.offsets + (long long)fl * (nseg + 1);
        const int o = 2 * m * p.ntypes + r;
        soff[k][r] = goff[o < nseg ? o : nseg];
    }
    __syncthreads();

    // producer side: thread 0 keeps kSfStages blocks in flight, across item boundaries
    ItemIter ahead{0, 0, 0, 0, 0, false};
    auto seek = [&](ItemIter& it) {
        // position on the first non-empty segment at or after (it.k, it.h, it.t); c0 = its start
        while (it.k < mine) {
            int fl, m, tile;
            item_pair(it.k, &fl, &m, &tile);
            while (it.h < 2) {
                if (2 * m + it.h < p.nz) {
                    while (it.t < p.ntypes) {
                        const int b = soff[it.k][it.h * p.ntypes + it.t], e = soff[it.k][it.h * p.ntypes + it.t + 1];
                        if (b < e) {
                            it.c0 = b;
                            it.e = e;
                            it.valid = true;
                            return;
                        }
                        ++it.t;
                    }
                }
                ++it.h;
                it.t = 0;
            }
            ++it.k;
            it.h = 0;
        }
        it.valid = false;
    };
    seek(ahead);
    int issued = 0;
    auto issue = [&]() {
        if (!ahead.valid) return;
        if (tid == 0) {
            int fl, m, tile;
            item_pair(ahead.k, &fl, &m, &tile);
            const int tile_x = tile / p.tiles_y, tile_y = tile - tile_x * p.tiles_y;
            const float2* tabx = p.tabx + ((long long)fl * p.tiles_x + tile_x) * p.cap * TX;
            const float2* taby = p.taby + ((long long)fl * p.tiles_y + tile_y) * p.cap * TY;
            const int nc = ahead.e - ahead.c0 < CH ? ahead.e - ahead.c0 : CH;
            cpx* dst = stage + (issued % kSfStages) * kStageElems;
            uint64_t* bar = &full[issued % kSfStages];
            mbar_expect_tx(bar, (uint32_t)(nc * (TX + TY) * sizeof(float2)));
            bulk_g2s(dst, tabx + (long long)ahead.c0 * TX, (uint32_t)(nc * TX * sizeof(float2)), bar);
            bulk_g2s(dst + CH * TX, taby + (long long)ahead.c0 * TY, (uint32_t)(nc * TY * sizeof(float2)), bar);
        }
        ++issued;
        ahead.c0 += CH;
        if (ahead.c0 >= ahead.e) {
            ++ahead.t;
            seek(ahead);
        }
    };
#pragma unroll 1
    for (int i = 0; i < kSfStages; ++i) issue();

    int used = 0;
#pragma unroll 1
    for (int k = 0; k < mine; ++k) {
        int fl, m, tile;
        item_pair(k, &fl, &m, &tile);
        const int ml = m - p.pair_begin;
        const int tile_x = tile / p.tiles_y, tile_y = tile - tile_x * p.tiles_y;
        const int kx0 = tile_x * TX, ky0 = tile_y * TY;
        const bool corner_tile = kx0 == 0 && ky0 == 0 && nq_x && nq_y;
        const float* snx = p.snx + (long long)fl * p.cap;
        const float* sny = p.sny + (long long)fl * p.cap;
        int gx[4], gy[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) gx[i] = kx0 + ty + 16 * i;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) gy[kk] = ky0 + tx + 16 * kk;

        float tot[4][2][4];     // [i][k][cc, ss, cs, sc], multiplied by the form factor
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                    for (int q = 0; q < 4; ++q) tot[i][kk][q] = 0.f;
            if (2 * m + h < p.nz) {
                for (int t = 0; t < p.ntypes; ++t) {
                    const int b = soff[k][h * p.ntypes + t], e = soff[k][h * p.ntypes + t + 1];
                    if (b == e) continue;
                    cpx P[4][2], Q[4][2];      // P = (cc, cs), Q = (sc, ss)
                    float corr = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            P[i][kk] = fast::c_make(0.f, 0.f);
                            Q[i][kk] = fast::c_make(0.f, 0.f);
                        }
                    for (int c0 = b; c0 < e; c0 += CH) {
                        const int nc = e - c0 < CH ? e - c0 : CH;
                        const cpx* ex = stage + (used % kSfStages) * kStageElems;
                        const cpx* ey = ex + CH * TX;
                        mbar_wait(&full[used % kSfStages], (uint32_t)((used / kSfStages) & 1));
#pragma unroll 4
                        for (int a = 0; a < nc; ++a) {
                            cpx ys[2];
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) ys[kk] = ey[a * TY + tx + 16 * kk];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const cpx x = ex[a * TX + ty + 16 * i];
                                const cpx xc = fast::c_make(fast::c_re(x), fast::c_re(x));
                                const cpx xs = fast::c_make(fast::c_im(x), fast::c_im(x));
#pragma unroll
                                for (int kk = 0; kk < 2; ++kk) {
                                    P[i][kk] = fast::fma2(xc, ys[kk], P[i][kk]);
                                    Q[i][kk] = fast::fma2(xs, ys[kk], Q[i][kk]);
                                }
                            }
                        }
                        if (corner_tile)
                            for (int a = 0; a < nc; ++a) corr += __ldg(&snx[c0 + a]) * __ldg(&sny[c0 + a]);
                        __syncthreads();             // every thread is done with this stage
                        ++used;
                        issue();                     // refill it with the block after the one already in flight
                    }
                    const float4* ff4 = p.ff4 + (long long)t * nsx * nsy;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            if (gx[i] < nsx && gy[kk] < nsy) {
                                const float4 f = __ldg(&ff4[gx[i] * nsy + gy[kk]]);
                                const float cc = fast::c_re(P[i][kk]), cs = fast::c_im(P[i][kk]);
                                const float sc = fast::c_re(Q[i][kk]);
                                float ss = fast::c_im(Q[i][kk]);
                                if (corner_tile && gx[i] == 0 && gy[kk] == 0) ss -= corr;
                                tot[i][kk][0] += cc * f.x;
                                tot[i][kk][1] += ss * f.y;
                                tot[i][kk][2] += cs * f.z;
                                tot[i][kk][3] += sc * f.w;
                            }
                        }
                }
            }
            if (h == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                        for (int q = 0; q < 4; ++q) stash[((i * 2 + kk) * 4 + q) * 256 + tid] = tot[i][kk][q];
            }
        }
        // tot = second slice (B), stash = first slice (A):  Z = S'_A + i*S'_B at up to four mirror positions
        float2* out = p.out + ((long long)fl * p.pair_count + ml) * p.nx * p.ny;
        auto emit = [&](int kx, int ky, float ar, float ai, float br, float bi) {
            out[kx * p.ny + ky] = make_float2(ar - bi, ai + br);      // nx*ny < 2^31
        };
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const int x = gx[i], y = gy[kk];
                if (x >= nsx || y >= nsy) continue;
                float A[4], B[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    A[q] = stash[((i * 2 + kk) * 4 + q) * 256 + tid];
                    B[q] = tot[i][kk][q];
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

// grow-only workspaces (phase tables, gathered form factors), one set per (device, stream): calls on different GPUs or
// on different streams of one GPU never share a block
struct SfWorkspace {
    void* ws = nullptr;
    size_t ws_bytes = 0;
    void* ff4 = nullptr;       // gathered form factors of the current psb_build_transmission call
    size_t ff4_bytes = 0;
};
std::mutex g_ws_mu;
std::map<std::pair<int, cudaStream_t>, SfWorkspace> g_ws;
SfWorkspace& workspace_of(cudaStream_t s) { return g_ws[std::make_pair(rt::device(), s)]; }    // g_ws_mu held

}  // namespace

void sf_fast_release() {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : g_ws) {
        cudaSetDevice(kv.first.first);
        rt::dev_free(kv.second.ws);
        rt::dev_free(kv.second.ff4);
    }
    cudaSetDevice(cur);
    g_ws.clear();
}

const float4* sf_fast_ff4(cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    return reinterpret_cast<const float4*>(workspace_of(s).ff4);
}

int sf_fast_prepare(const float* ff, int ntypes, int nx, int ny, cudaStream_t s, cudaStream_t owner) {
    const size_t n = (size_t)ntypes * StructureFactorPaired::slots(nx) * StructureFactorPaired::slots(ny);
    float4* ff4 = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        SfWorkspace& w = workspace_of(owner);
        if (n * sizeof(float4) > w.ff4_bytes) {
            cudaError_t e = cudaStreamSynchronize(owner);
            if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("sf workspace sync: ") + cudaGetErrorString(e));
            graph_cache_release();                         // recorded launch sequences point at the block freed here
            rt::dev_free(w.ff4);
            w.ff4 = rt::dev_alloc(n * sizeof(float4));
            w.ff4_bytes = w.ff4 ? n * sizeof(float4) : 0;
            if (!w.ff4) return PSB_ERR_NOMEM;
        }
        ff4 = reinterpret_cast<float4*>(w.ff4);
    }
    const long long blocks = (long long)((n + 255) / 256);
    ff4_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, s>>>(ff, ff4, ntypes, nx, ny);
    ++launch_counter();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("ff4 launch: ") + cudaGetErrorString(e));
    return PSB_OK;
}

bool sf_fast_supported(int ntypes) { return ntypes <= kMaxStagedTypes; }

int launch_sf_fast(const int* offsets, const unsigned int* ux, const unsigned int* uy, int cap, int nz, int ntypes, int nx,
                   int ny, int pair_begin, int pair_count, int nf, float2* out, cudaStream_t s, cudaStream_t owner) {
    SfFastParams p;
    std::memset(&p, 0, sizeof(p));
    p.offsets = offsets; p.ux = ux; p.uy = uy; p.cap = cap; p.nz = nz; p.ntypes = ntypes; p.nx = nx; p.ny = ny;
    p.pair_begin = pair_begin; p.pair_count = pair_count; p.out = out;
    if (ntypes > kMaxStagedTypes) return fail(PSB_ERR_UNSUPPORTED, "pipelined structure factor: more than 64 atom types");
    p.tiles_x = (StructureFactorPaired::slots(nx) + TX - 1) / TX;
    p.tiles_y = (StructureFactorPaired::slots(ny) + TY - 1) / TY;
    const size_t nx_elems = (size_t)nf * p.tiles_x * cap * TX, ny_elems = (size_t)nf * p.tiles_y * cap * TY;
    const size_t sn_elems = (size_t)nf * cap;
    const size_t need = (nx_elems + ny_elems) * sizeof(float2) + 2 * sn_elems * sizeof(float) + 256;
    {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        SfWorkspace& w = workspace_of(owner);
        if (need > w.ws_bytes) {
            cudaError_t e = cudaStreamSynchronize(owner);  // kernels of earlier chunks may still read the old block
            if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("sf workspace sync: ") + cudaGetErrorString(e));
            graph_cache_release();                         // recorded launch sequences point at the block freed here
            rt::dev_free(w.ws);
            w.ws = rt::dev_alloc(need);
            w.ws_bytes = w.ws ? need : 0;
            if (!w.ws) return PSB_ERR_NOMEM;
        }
        if (!w.ff4) return fail(PSB_ERR_INVALID, "launch_sf_fast without sf_fast_prepare");
        p.ff4 = reinterpret_cast<const float4*>(w.ff4);
        p.tabx = reinterpret_cast<float2*>(w.ws);
        p.taby = p.tabx + nx_elems;
        p.snx = reinterpret_cast<float*>(p.taby + ny_elems);
        p.sny = p.snx + sn_elems;
    }
    static rt::PerDeviceOnce once;
    int rc0 = once.run([] {
        cudaError_t e = cudaFuncSetAttribute(sf_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSfSmem);
        if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("sf tiles: ") + cudaGetErrorString(e));
        return (int)PSB_OK;
    });
    if (rc0 != PSB_OK) return rc0;
    if (cap > 0) {
        cudaError_t e1 = pdl_launch(phase_tables_kernel, dim3((cap + 7) / 8, nf), dim3(256), 0, s, p);
        ++launch_counter();
        if (e1 != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("phase tables launch: ") + cudaGetErrorString(e1));
    }
    cudaError_t e = cudaSuccess;
    if (PSB_SF_PERSISTENT) {
        static rt::PerDeviceOnce once2;
        int rc1 = once2.run([] {
            cudaError_t e2 = cudaFuncSetAttribute(sf_tiles_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSfSmem);
            if (e2 != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("sf tiles: ") + cudaGetErrorString(e2));
            return (int)PSB_OK;
        });
        if (rc1 != PSB_OK) return rc1;
        // two CTAs per SM; a grid that is a multiple of the tile count keeps every CTA on one tile
        const int tiles = p.tiles_x * p.tiles_y;
        const int slots = 2 * rt::sm_count();
        const int grid_full = tiles <= slots ? slots / tiles * tiles : slots;
        // pair images per launch: at most kSfMaxItems items per CTA
        const long long imgs_total = (long long)pair_count * nf;
        const long long imgs_per_launch = std::max<long long>(1, (long long)kSfMaxItems * grid_full / tiles);
        if (imgs_total <= imgs_per_launch) {
            const long long n_items = imgs_total * tiles;
            const int grid = (int)std::min<long long>(grid_full, n_items);
            e = pdl_launch(sf_tiles_persistent_kernel, dim3(grid), dim3(256), kSfSmem, s, p, (int)n_items);
        } else {
            // a chunk of more images than that (tiny grids): launch by ranges of pairs of one frame at a time
            for (int f = 0; f < nf && e == cudaSuccess; ++f)
                for (long long m0 = 0; m0 < pair_count && e == cudaSuccess; m0 += imgs_per_launch) {
                    SfFastParams q = p;
                    const int cnt = (int)std::min<long long>(imgs_per_launch, pair_count - m0);
                    q.offsets = p.offsets + (long long)f * (nz * ntypes + 1);
                    q.ux = p.ux + (long long)f * cap; q.uy = p.uy + (long long)f * cap;
                    q.tabx = p.tabx + (long long)f * p.tiles_x * cap * TX; q.taby = p.taby + (long long)f * p.tiles_y * cap * TY;
                    q.snx = p.snx + (long long)f * cap; q.sny = p.sny + (long long)f * cap;
                    q.pair_begin = pair_begin + (int)m0; q.pair_count = cnt;
                    q.out = out + ((long long)f * pair_count + m0) * nx * ny;
                    const long long n_items = (long long)cnt * tiles;
                    e = pdl_launch(sf_tiles_persistent_kernel, dim3((int)std::min<long long>(grid_full, n_items)), dim3(256), kSfSmem, s, q, (int)n_items);
                    ++launch_counter();
                }
            --launch_counter();
        }
    } else {
        e = pdl_launch(sf_tiles_kernel, dim3(p.tiles_x * p.tiles_y, pair_count, nf), dim3(256), kSfSmem, s, p);
    }
    if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("sf fast launch: ") + cudaGetErrorString(e));
    return PSB_OK;
}

}  // namespace psb
