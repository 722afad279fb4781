// Structure-factor sum of DENSE slices through a one-dimensional non-uniform FFT (reference: src/multislice/potentials.py
// :319-330, the einsum over exp(-2 pi i kx x) exp(-2 pi i ky y) of the atoms of a slice, times the form factor).
//
// The direct sum (sf_fast.cu) costs 4 * atoms * (nx/2) * (ny/2) FMAs per slice pair: 0.65 GFMA at config C4 (1024 x 1024,
// ~620 atoms per pair), 78 % of that configuration's potential build even at 59 % of the fp32 peak.  Here the x axis is
// handled the NUFFT way (type 1, Dutt-Rokhlin / Barnett's "exponential of semicircle" kernel) and the y axis stays exact:
//
//     G[j, ky] = sum_a phi(M u_a - j) exp(-2 pi i ky y_a)          j = 0 .. M-1,  M = 2 nx fine cells, 8 taps per atom
//     S[kx, ky] = f(kx, ky) / phi_hat(kx / M) * FFT_j(G)[kx, ky]    |kx| <= nx/2
//
// with phi(z) = exp(beta (sqrt(1 - (z/4)^2) - 1)), beta = 2.3 * 8: the aliasing error of this kernel at twofold
// oversampling is ~1e-7 of the spectrum (measured against the direct sum: tests/test_gpu_parity.py), i.e. the float32
// round-off of the sum itself.  Work per slice pair: atoms * 8 taps * ny complex FMAs + one 2 nx-point FFT per column and
// atom type -- 8 x fewer flops than the direct sum at C4 and none of them atom-count dependent beyond the spreading.
//
//   K1  nufft_prep_kernel   one CTA per (slice pair, frame): orders the pair's atoms by (type, x bin of 8 fine cells) --
//                           deterministically: ties keep list order -- and sums the (Nyquist, Nyquist) corner term.
//   K2  nufft_cols_kernel   persistent, one CTA per SM walking (W adjacent ky columns, slice pair, frame) tiles: builds the G tile in shared memory
//                           (thread (column, bin) spreads the atoms of its bin; even bins, then odd bins, so no two threads
//                           touch a cell at the same time and the sums are reproducible), transforms it in place
//                           (fast_fft.cuh: radix 16 x 8 x 16 for M = 2048, 16 x 4 x 16 for M = 1024), keeps the nx low
//                           frequencies, divides out phi_hat, applies the form factor, accumulates over atom types and
//                           writes the slice-pair spectrum S_2m + i S_2m+1 where the inverse transforms expect it.
//
// Hermitian part (the reference's Re(ifft2(.)), potentials.py:336-337): the ky Nyquist column is spread with cos(pi ny y)
// instead of the complex phase, the kx Nyquist row is the mean of the +nx/2 and -nx/2 outputs of the fine transform, and
// the corner gets the -sum sin sin term from K1 -- the same spectrum, entry for entry, as the direct kernels build.
#include "fast_fft.cuh"
#include "fast_path.h"
#include "graph_cache.h"
#include "pdl.cuh"
#include "potential_kernels.cuh"
#include "psb_rt.h"
#include "tables.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace psb {

namespace {

using fast::cpx;

constexpr int kTaps = 8;
// x bins the records are ordered by: kBinCells fine cells each.  The gather's warp of 32 rows needs the records of cells
// [r - 4, r + 35]: six bins of 8 cells = 48 cells, or ten bins of 4 = exactly those 40 (17 % fewer records walked)
// rows per lane of the gather: 2 shares a record's fetches between two rows 32 apart and narrows the warp's window, but
// measured slower (11.65 against 10.65 ms per 8 frames of C4: registers, profiles/r2ar_nufft_rows_per_lane.txt)
#ifndef PSB_NUFFT_ROWS_PER_LANE
#define PSB_NUFFT_ROWS_PER_LANE 1
#endif
#ifndef PSB_NUFFT_BIN_SHIFT
#define PSB_NUFFT_BIN_SHIFT 2
#endif
constexpr int kBinShift = PSB_NUFFT_BIN_SHIFT, kBinCells = 1 << kBinShift;
constexpr float kBeta = 2.30f * kTaps;
constexpr int kMaxKeys = 8192;          // ntypes * bins the ordering kernel can histogram in shared memory

std::atomic<int> g_sf_mode{0};          // 0 auto, 1 direct sum always, 2 NUFFT wherever it is supported

struct NufftParams {
    const int* offsets;         // (F, nseg+1), first frame of the CALL (K1 covers every frame and pair of it at once)
    const unsigned int* ux;     // (F, cap) fixed-point fractions, grouped by (slice, type) segment
    const unsigned int* uy;
    int cap, nz, ntypes, nx, ny;
    int npairs;                 // slice pairs per frame
    int frame0, pair_begin, pair_count;     // K2: this chunk covers frames [frame0, frame0 + nf) x pairs [pair_begin, pair_begin + pair_count)
    int nb, log_m;              // x bins per pair (= M / kBinCells), log2(M)
    // written by K1, read by K2
    int* xoff;                  // (F, npairs, ntypes*nb + 1) offsets into the pair's record range, by (type, bin)
    unsigned int* rx;           // (F, cap) records of a pair, ordered by (type, bin), list order inside
    unsigned int* ry;
    unsigned int* rpar;         // 0: first slice of the pair (real part), 1: second (imaginary part)
    float* rwt;                 // (F, cap, 8) the record's eight tap weights phi(k - 3 - frac), k = 0..7 (tap rows: cell - 3 + k)
    float* corner;              // (F, npairs, ntypes, 2)  sum sin(pi nx u) sin(pi ny v) per type and slice of the pair
    const float* ff;            // (ntypes, nx, ny) form factors
    const float* dec;           // (nx) 1 / phi_hat(kx / M)
    const float2* tw;           // staged twiddles of the M-point plan
    float2* out;                // (nf, pair_count, nx, ny)
};

// phi(z) = exp(beta (sqrt(1 - z^2/16) - 1)), |z| <= 4, through the SFU with one correction step each (relative error
// ~2e-7; the library sqrtf / expf cost 5 x the instructions, and a tile evaluates ~40 000 weights)
__device__ __forceinline__ float es_weight(float z) {
    const float s = fmaxf(fmaf(-z * z, 1.0f / 16.0f, 1.0f), 1e-30f);
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    float q = s * r;                                             // ~sqrt(s)
    q = fmaf(0.5f * r, fmaf(-q, q, s), q);                       // one Newton step
    const float x = kBeta * (q - 1.0f);                          // in [-beta, 0]
    const float kL2e = 1.4426950408889634f, kL2eLo = 1.925963033500e-8f;
    const float t = x * kL2e;
    const float rem = fmaf(x, kL2e, -t) + x * kL2eLo;            // what the rounding of t lost, in units of log2
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return fmaf(e, rem * 0.6931471805599453f, e);
}

// exp(-2*pi*i*m*v) with v a 32-bit turn fraction: exact integer phase reduction, then the SFU on an angle in [-pi, pi)
// (absolute error ~4e-7, the accuracy of the direct kernels' phase tables, sf_fast.cu)
__device__ __forceinline__ float2 unit_phase_fast(int m, unsigned int v) {
    const int ph = (int)((unsigned int)m * v);
    float sn, cs;
    __sincosf((float)ph * (3.14159265358979323846f * 4.656612873077393e-10f), &sn, &cs);
    return make_float2(cs, -sn);
}

// ---- K1 ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nufft_prep_kernel(const NufftParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* hist = reinterpret_cast<int*>(smem_raw);                      // [nkeys + 1]
    const int nkeys = p.ntypes * p.nb;
    unsigned short* keys = reinterpret_cast<unsigned short*>(hist + nkeys + 1 + 256);      // [kChunk]
    int* part = hist + nkeys + 1;                                      // [256] scan partials
    double* red = reinterpret_cast<double*>(smem_raw);                 // reduction scratch (reused after the ordering)
    constexpr int kChunk = 4096;
    const int t = threadIdx.x;
    const int m = blockIdx.x, f = blockIdx.y;                          // every pair of every frame of the call
    const int nseg = p.nz * p.ntypes;
    const int* off = p.offsets + (long long)f * (nseg + 1);
    const int s0 = 2 * m, s1 = (2 * m + 2 < p.nz) ? 2 * m + 2 : p.nz;
    const int begin = off[s0 * p.ntypes], end = off[s1 * p.ntypes], n = end - begin;
    const unsigned int* ux = p.ux + (long long)f * p.cap;
    const unsigned int* uy = p.uy + (long long)f * p.cap;
    const int shift = 32 - p.log_m + kBinShift;                        // bin = fine cell / kBinCells

    auto seg_of = [&](int i) {                                         // list index -> segment (2 * ntypes candidates)
        int seg = s0 * p.ntypes;
        while (seg + 1 < s1 * p.ntypes && off[seg + 1] <= i) ++seg;
        return seg;
    };
    auto key_of = [&](int i, int* par) {
        const int seg = seg_of(i);
        *par = seg / p.ntypes - s0;
        return (seg % p.ntypes) * p.nb + (int)(ux[i] >> shift);
    };

    for (int k = t; k <= nkeys; k += 256) hist[k] = 0;
    __syncthreads();
    for (int i = begin + t; i < end; i += 256) {
        int par;
        atomicAdd(&hist[key_of(i, &par) + 1], 1);                      // integer counts: order-independent
    }
    __syncthreads();
    // exclusive scan of hist[1..nkeys] in place (hist[k] = atoms with a smaller key)
    {
        const int chunk = (nkeys + 255) / 256;
        const int i0 = 1 + t * chunk, i1 = (i0 + chunk < nkeys + 1) ? i0 + chunk : nkeys + 1;
        int s = 0;
        for (int i = i0; i < i1; ++i) s += hist[i];
        part[t] = s;
        __syncthreads();
        int base = 0;
        for (int k = 0; k < t; ++k) base += part[k];
        for (int i = i0; i < i1; ++i) {
            base += hist[i];
            hist[i] = base;
        }
        __syncthreads();
    }
    int* xoff = p.xoff + ((long long)f * p.npairs + m) * (nkeys + 1);
    for (int k = t; k <= nkeys; k += 256) xoff[k] = hist[k];
    // position of atom i = hist[key] + number of earlier list entries with the same key (stable, hence reproducible)
    unsigned int* rx = p.rx + (long long)f * p.cap + begin;
    unsigned int* ry = p.ry + (long long)f * p.cap + begin;
    unsigned int* rpar = p.rpar + (long long)f * p.cap + begin;
    for (int a0 = 0; a0 < n; a0 += 256) {
        const int i = a0 + t;
        int par = 0, same = 0;
        const int key = i < n ? key_of(begin + i, &par) : -1;
        for (int c0 = 0; c0 < a0 + 256 && c0 < n; c0 += kChunk) {
            const int mchunk = n - c0 < kChunk ? n - c0 : kChunk;
            __syncthreads();
            for (int k = t; k < mchunk; k += 256) {
                int pk;
                keys[k] = (unsigned short)key_of(begin + c0 + k, &pk);
            }
            __syncthreads();
            const int lim = (i - c0) < mchunk ? (i - c0) : mchunk;       // entries of this chunk that precede i
            for (int k = 0; k < lim; ++k) same += keys[k] == key;
        }
        if (i < n) {
            const int pos = hist[key] + same;
            const unsigned int u = ux[begin + i];
            rx[pos] = u;
            ry[pos] = uy[begin + i];
            rpar[pos] = (unsigned int)par;
            const unsigned int frac_bits = 32 - p.log_m;
            const float fr = (float)(u & ((1u << frac_bits) - 1u)) * (1.0f / (float)(1u << frac_bits));
            float4* wq = reinterpret_cast<float4*>(p.rwt + ((long long)f * p.cap + begin + pos) * kTaps);
            wq[0] = make_float4(es_weight(-3.f - fr), es_weight(-2.f - fr), es_weight(-1.f - fr), es_weight(0.f - fr));
            wq[1] = make_float4(es_weight(1.f - fr), es_weight(2.f - fr), es_weight(3.f - fr), es_weight(4.f - fr));
        }
    }
    // corner term per (type, slice of the pair): sum sin(pi nx u) sin(pi ny v), fixed summation order
    __syncthreads();
    for (int c = 0; c < 2 * p.ntypes; ++c) {
        const int typ = c >> 1, par = c & 1;
        const int seg = (s0 + par) * p.ntypes + typ;
        double acc = 0.0;
        if (s0 + par < p.nz)
            for (int i = off[seg] + t; i < off[seg + 1]; i += 256) {
                const float2 ex = unit_phase(p.nx / 2, ux[i]), ey = unit_phase(p.ny / 2, uy[i]);     // (cos, -sin)(pi n u)
                acc += (double)(ex.y * ey.y);
            }
        red[t] = acc;
        __syncthreads();
        for (int h = 128; h > 0; h >>= 1) {
            if (t < h) red[t] += red[t + h];
            __syncthreads();
        }
        if (t == 0) p.corner[(((long long)f * p.npairs + m) * p.ntypes + typ) * 2 + par] = (float)red[0];
        __syncthreads();
    }
}

// ---- K2 ------------------------------------------------------------------------------------------------------
template <int M>
struct NufftCfg {
    static constexpr int T = M / 16;                 // threads per column = x bins per pair
    static constexpr int kThreads = 512;
    static constexpr int W = kThreads / T;           // columns per tile: 4 (M = 2048), 8 (M = 1024)
    static constexpr int kRows = M + M / 16;         // padded
    static constexpr size_t kTileBytes = (size_t)kRows * W * sizeof(float2);
    static constexpr int kStage = W == 4 ? 1536 : 1024;      // records of one atom type staged in shared memory (more: read from L2)
    static constexpr size_t kEOffset = (kTileBytes + 3 * kStage * sizeof(unsigned int) + (M / kBinCells + 1) * sizeof(int) + (M / 2) * sizeof(float) + 15) / 16 * 16;
    static constexpr size_t kSmem = kEOffset + (size_t)kStage * W * sizeof(float2) + (size_t)kStage * 8 * sizeof(float);
    static constexpr int kLogM = M == 2048 ? 11 : 10;
};

template <int M>
struct NufftXchg {
    cpx* t;
    int c;
    __device__ __forceinline__ cpx* buf(int) const { return t; }
    __device__ __forceinline__ int at(int q) const { return (q + (q >> 4)) * NufftCfg<M>::W + c; }
    __device__ __forceinline__ void after_store(int) const { __syncthreads(); }
    __device__ __forceinline__ void after_load(int) const { __syncthreads(); }
    __device__ __forceinline__ void mid_sync(int) const { __syncthreads(); }
};

// two adjacent packed complex values with one 128-bit load (p 16-byte aligned)
__device__ __forceinline__ void load_pair(const cpx* p, cpx& a, cpx& b) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
    a = v.x;
    b = v.y;
#else
    a = p[0];
    b = p[1];
#endif
}

template <int M>
__global__ void __launch_bounds__(NufftCfg<M>::kThreads, 1) nufft_cols_kernel(const NufftParams p, const int n_tiles) {
    using C = NufftCfg<M>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tile = reinterpret_cast<float2*>(smem_raw);
    unsigned int* s_rx = reinterpret_cast<unsigned int*>(smem_raw + C::kTileBytes);      // [kStage] records of the current (pair, type)
    unsigned int* s_ry = s_rx + C::kStage;
    unsigned int* s_rp = s_ry + C::kStage;
    int* s_xoff = reinterpret_cast<int*>(s_rp + C::kStage);                              // [M / kBinCells + 1]
    float* s_dec = reinterpret_cast<float*>(s_xoff + M / kBinCells + 1);                 // [nx]
    float2* s_e = reinterpret_cast<float2*>(smem_raw + C::kEOffset);                     // [kStage][W] exp(-2 pi i ky y) per record and column
    float* s_w = reinterpret_cast<float*>(s_e + (size_t)C::kStage * C::W);               // [kStage][8] tap weights
    const int tid = threadIdx.x, c = tid % C::W, j = tid / C::W;
    constexpr int nx = M / 2;
    // persistent CTA: twiddles and the deconvolution table are loaded once, a contiguous range of (frame, pair, column
    // tile) triples is walked pair by pair so a pair's records are staged once for all of its column tiles
    fast::Twiddles<M> tw;
    tw.load(p.tw, j);
    for (int i = tid; i < nx; i += C::kThreads) s_dec[i] = p.dec[i];
    const NufftXchg<M> xc{reinterpret_cast<cpx*>(tile), c};
    const int nkeys = p.ntypes * (M / kBinCells);
    const int nseg = p.nz * p.ntypes;
    const int tiles_per_pair = p.ny / C::W;
    constexpr unsigned int kFracBits = 32 - C::kLogM;
    constexpr float kFracScale = 1.0f / (float)(1u << kFracBits);
    // Spreading: warp w owns the fine rows [RW*w, RW*(w+1)) of the tile, 32 at a time (see the gather below)
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int RW = M / 16;                                    // rows per warp
    constexpr int kLogRW = C::kLogM - 4;
    constexpr int kBins = M / kBinCells;                          // x bins (the records are ordered by them)
    const int tile0 = (int)((long long)n_tiles * blockIdx.x / gridDim.x);
    const int tile1 = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
    int staged_img = -1, cur_img = -1, type_r0 = 0, type_r1 = 0;
    const int* xoff = nullptr;
    const unsigned int *rx = nullptr, *ry = nullptr, *rpar = nullptr;
    const float *rwt = nullptr, *corner = nullptr;

    for (int tix = tile0; tix < tile1; ++tix) {
        const int img = tix / tiles_per_pair;                    // frame * pair_count + pair-in-chunk
        const int my = (tix - img * tiles_per_pair) * C::W + c;  // column, fft order
        const int msy = my < (p.ny + 1) / 2 ? my : my - p.ny;
        const bool nyq_y = (p.ny % 2 == 0) && my == p.ny / 2;
        if (img != cur_img) {                                    // per-pair pointers: a chain of dependent global loads, once per pair
            cur_img = img;
            const int fl = img / p.pair_count, ml = img - fl * p.pair_count;
            const int f = p.frame0 + fl, m = p.pair_begin + ml;
            xoff = p.xoff + ((long long)f * p.npairs + m) * (nkeys + 1);
            const long long rbase = (long long)f * p.cap + p.offsets[(long long)f * (nseg + 1) + 2 * m * p.ntypes];
            rx = p.rx + rbase;
            ry = p.ry + rbase;
            rpar = p.rpar + rbase;
            rwt = p.rwt + rbase * kTaps;
            corner = p.corner + ((long long)f * p.npairs + m) * p.ntypes * 2;
            type_r0 = xoff[0];                                   // record range of type 0 (the only one in single-type runs)
            type_r1 = xoff[kBins];
        }
        cpx acc_out[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) acc_out[s] = fast::c_make(0.f, 0.f);

        for (int z = 0; z < p.ntypes; ++z) {
            // form factors of this tile's 8 rows per thread: issued now, needed after the transform
            float g8[8];
            {
                const float* ffz = p.ff + ((long long)z * nx) * p.ny + my;
#pragma unroll
                for (int s = 0; s < 8; ++s) g8[s] = __ldg(ffz + (long long)(j + C::T * s) * p.ny);
            }
            // ---- stage this (pair, type)'s records and bin offsets (kept across the pair's column tiles when there is
            //      one atom type), clear the tile
            const int r0 = z == 0 ? type_r0 : xoff[z * kBins], r1 = z == 0 ? type_r1 : xoff[(z + 1) * kBins];
            if (p.ntypes > 1 || staged_img != img) {
                const int nstage = (r1 - r0) < C::kStage ? (r1 - r0) : C::kStage;
                for (int i = tid; i < nstage; i += C::kThreads) {
                    s_rx[i] = rx[r0 + i];
                    s_ry[i] = ry[r0 + i];
                    s_rp[i] = rpar[r0 + i];
                }
                for (int i = tid; i < nstage * 2; i += C::kThreads)
                    reinterpret_cast<float4*>(s_w)[i] = reinterpret_cast<const float4*>(rwt + (long long)r0 * kTaps)[i];
                for (int i = tid; i <= kBins; i += C::kThreads) s_xoff[i] = xoff[z * kBins + i] - r0;
                staged_img = img;
            }
            {   // phase factors of the staged records for this tile's columns (a record is used by up to two warps and eight taps)
                const int nstage = (r1 - r0) < C::kStage ? (r1 - r0) : C::kStage;
                __syncthreads();                                 // the staged records are visible
                for (int q = tid; q < nstage * C::W; q += C::kThreads) {
                    const int i = q / C::W, cc = q - i * C::W;
                    const int myc = my - c + cc;
                    const int msc = myc < (p.ny + 1) / 2 ? myc : myc - p.ny;
                    float2 e = ((p.ny % 2 == 0) && myc == p.ny / 2) ? make_float2(unit_phase_fast(p.ny / 2, s_ry[i]).x, 0.f)   // cos(pi ny v): the Hermitian part
                                                                     : unit_phase_fast(msc, s_ry[i]);
                    if (s_rp[i]) e = make_float2(-e.y, e.x);                                      // second slice of the pair: times i
                    s_e[q] = e;
                }
            }
            __syncthreads();                                     // phase factors staged
            // ---- spread as a GATHER: lane = fine row, the warp's 32 consecutive rows walk the records whose taps can
            //      reach them (bins of 8 cells around the rows; every lane reads the same record: broadcast loads), each
            //      lane adds the tap that lands on its row -- w[record][row - cell + 3] -- times the record's phase factors
            //      to its W column accumulators and stores the finished cells once.  No read-modify-write of shared
            //      memory, no ordering between records beyond program order, no barrier: the earlier scatter forms were
            //      bound by exactly those (profiles/r2_nufft_spread_history.txt).
            // A lane takes RPL rows 32 apart (1 by default; with 2 the record's cell, tap index -- k and k + 32 pick the same
            // weight slot -- and phase factors are fetched once for both rows, and the warp's window is 64 + 8 cells instead of
            // 2 x (32 + 8))
            constexpr int RPL = (C::W == 4 && PSB_NUFFT_ROWS_PER_LANE == 2) ? 2 : 1;
#pragma unroll 1
            for (int pass = 0; pass < RW / (32 * RPL); ++pass) {
                const int r_first = RW * warp + 32 * RPL * pass;
                const int r = r_first + lane;
                cpx acc[RPL][C::W];                              // packed: one FFMA2 per record, row and column (weight broadcast to both halves)
#pragma unroll
                for (int rr = 0; rr < RPL; ++rr)
#pragma unroll
                    for (int cc = 0; cc < C::W; ++cc) acc[rr][cc] = fast::c_make(0.f, 0.f);
                const cpx* s_e2 = reinterpret_cast<const cpx*>(s_e);
                // weights of one record for the lane's rows: the tap that lands on the row, or 0
                auto taps = [&](unsigned int rxv, const float* wrec, float* w) {
                    const int cell = (int)(rxv >> kFracBits);
                    const int k0 = ((r - cell + 3 + M / 2) & (M - 1)) - M / 2;           // periodic distance, in taps
                    const float ww = wrec[k0 & (kTaps - 1)];
                    w[0] = (unsigned int)k0 < (unsigned int)kTaps ? ww : 0.f;
                    if (RPL == 2) {
                        const int k1 = ((r + 32 - cell + 3 + M / 2) & (M - 1)) - M / 2;  // same slot: 32 is a multiple of kTaps
                        w[RPL - 1] = (unsigned int)k1 < (unsigned int)kTaps ? ww : 0.f;
                    }
                };
                auto add_record = [&](int i, const float* w) {
                    // the record's W phase factors are contiguous: 128-bit broadcast loads, two columns each (ncu r2ak: the
                    // kernel's busiest unit is the L1 / shared-memory pipe at 73 %, and this loop issues most of its loads)
#pragma unroll
                    for (int cc = 0; cc < C::W; cc += 2) {
                        cpx e0, e1;
                        load_pair(s_e2 + i * C::W + cc, e0, e1);
#pragma unroll
                        for (int rr = 0; rr < RPL; ++rr) {
                            const cpx w2 = fast::c_make(w[rr], w[rr]);
                            acc[rr][cc] = fast::fma2(w2, e0, acc[rr][cc]);
                            acc[rr][cc + 1] = fast::fma2(w2, e1, acc[rr][cc + 1]);
                        }
                    }
                };
                // staged records: branch-free body, four records per trip (their loads are independent, so the shared-memory
                // latency is paid once per four; a record whose taps miss this lane's rows contributes weight 0)
                auto staged4 = [&](int i) {
                    float w[4][RPL];
#pragma unroll
                    for (int q = 0; q < 4; ++q) taps(s_rx[i + q], s_w + (i + q) * kTaps, w[q]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) add_record(i + q, w[q]);
                };
                auto staged1 = [&](int i) {
                    float w[RPL];
                    taps(s_rx[i], s_w + i * kTaps, w);
                    add_record(i, w);
                };
                auto unstaged = [&](int i) {                     // beyond the staging area: from L2, phase factors on the fly
                    float w[RPL];
                    taps(rx[r0 + i], rwt + (long long)(r0 + i) * kTaps, w);
                    bool any = false;
#pragma unroll
                    for (int rr = 0; rr < RPL; ++rr) any = any || w[rr] != 0.f;
                    if (any) {
                        const unsigned int v = ry[r0 + i];
                        const bool par = rpar[r0 + i] != 0;
#pragma unroll
                        for (int cc = 0; cc < C::W; ++cc) {
                            const int myc = my - c + cc;
                            const int msc = myc < (p.ny + 1) / 2 ? myc : myc - p.ny;
                            float2 e = ((p.ny % 2 == 0) && myc == p.ny / 2) ? make_float2(unit_phase_fast(p.ny / 2, v).x, 0.f)
                                                                             : unit_phase_fast(msc, v);
                            if (par) e = make_float2(-e.y, e.x);
#pragma unroll
                            for (int rr = 0; rr < RPL; ++rr)
                                acc[rr][cc] = fast::fma2(fast::c_make(w[rr], w[rr]), fast::c_make(e.x, e.y), acc[rr][cc]);
                        }
                    }
                };
                auto walk = [&](int i0, int i1) {
                    const int i1s = i1 < C::kStage ? i1 : C::kStage;
                    int i = i0;
                    for (; i + 4 <= i1s; i += 4) staged4(i);
                    for (; i < i1s; ++i) staged1(i);
                    for (; i < i1; ++i) unstaged(i);
                };
                // records with cell in [r_first - 4, r_first + 32 RPL + 3]
                const int b_lo = (r_first - 4) >> kBinShift, b_hi = (r_first + 32 * RPL + 3) >> kBinShift;      // -4 >> s = -1
                if (b_lo < 0) walk(s_xoff[kBins - 1], s_xoff[kBins]);                 // wraps around the periodic axis
                walk(s_xoff[b_lo < 0 ? 0 : b_lo], s_xoff[(b_hi > kBins - 1 ? kBins - 1 : b_hi) + 1]);
                if (b_hi > kBins - 1) walk(s_xoff[0], s_xoff[1]);
#pragma unroll
                for (int rr = 0; rr < RPL; ++rr) {
                    const int rw = r + 32 * rr;
                    cpx* cells = reinterpret_cast<cpx*>(tile) + (rw + (rw >> 4)) * C::W;
#pragma unroll
                    for (int cc = 0; cc < C::W; ++cc) cells[cc] = acc[rr][cc];
                }
            }
            __syncthreads();
            // ---- M-point transform along x, in place in the tile
            cpx v16[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) v16[e] = reinterpret_cast<const cpx*>(tile)[xc.at(j + e * C::T)];
            __syncthreads();
            cpx nyq_hold = fast::c_make(0.f, 0.f);
            fast::line_fft<M, -1>(
                [&](int e) { return v16[e]; },
                [&](int t, cpx a) {
                    // fine frequency j + T*t; kept: t in {0..3} (kx >= 0) and {12..15} (kx < 0), coarse row kxi = j + T*s
                    if (t >= 4 && t < 12) {
                        if (t == 4) nyq_hold = a;                    // +nx/2 (j == 0 only)
                        return;
                    }
                    const int s = t < 4 ? t : t - 8;
                    const float dk = s_dec[j + C::T * s];
                    if (s == 4 && j == 0) {                          // kx Nyquist row: mean of the +nx/2 and -nx/2 outputs
                        a = fast::mul2(fast::add2(a, nyq_hold), fast::c_make(0.5f, 0.5f));
                        if (nyq_y)                                    // corner: cos cos - sin sin
                            a = fast::sub2(a, fast::c_make(corner[2 * z] / dk, corner[2 * z + 1] / dk));
                    }
                    const float g = g8[s < 8 ? s : 0] * dk;
                    acc_out[s < 8 ? s : 0] = fast::fma2(a, fast::c_make(g, g), acc_out[s < 8 ? s : 0]);
                },
                tw, j, xc, 0);
        }
        float2* out = p.out + ((long long)img * nx) * p.ny + my;
#pragma unroll
        for (int s = 0; s < 8; ++s) out[(long long)(j + C::T * s) * p.ny] = make_float2(fast::c_re(acc_out[s]), fast::c_im(acc_out[s]));
    }
}

// ---- host side -------------------------------------------------------------------------------------------
struct NufftWorkspace {
    void* block = nullptr;
    size_t bytes = 0;
};
std::mutex g_mu;
std::map<std::pair<int, cudaStream_t>, NufftWorkspace> g_ws;      // (device, owner stream)
std::map<std::pair<int, int>, float*> g_dec;                       // (device, nx) -> 1 / phi_hat

// phi_hat(xi) = integral_{-4}^{4} phi(z) cos(2 pi xi z) dz, composite Simpson (phi is smooth inside, ~1e-8 at the ends)
double phi_hat(double xi) {
    const int n = 4096;
    const double a = -0.5 * kTaps, h = (double)kTaps / n, pi = 3.14159265358979323846;
    auto fn = [&](double z) {
        const double s = 1.0 - z * z / 16.0;
        return std::exp((double)kBeta * (std::sqrt(s > 0 ? s : 0) - 1.0)) * std::cos(2.0 * pi * xi * z);
    };
    double acc = fn(a) + fn(a + n * h);
    for (int i = 1; i < n; ++i) acc += fn(a + i * h) * ((i & 1) ? 4.0 : 2.0);
    return acc * h / 3.0;
}

int dec_table(int nx, const float** out, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_mu);
    const auto key = std::make_pair(rt::device(), nx);
    auto it = g_dec.find(key);
    if (it != g_dec.end()) {
        *out = it->second;
        return PSB_OK;
    }
    std::vector<float> host(nx);
    const int M = 2 * nx;
    for (int i = 0; i < nx; ++i) {
        const int mfreq = i < (nx + 1) / 2 ? i : i - nx;
        host[i] = (float)(1.0 / phi_hat((double)mfreq / (double)M));
    }
    float* d = static_cast<float*>(rt::dev_alloc(nx * sizeof(float)));
    if (!d) return PSB_ERR_NOMEM;
    int rc = rt::h2d(d, host.data(), nx * sizeof(float), s);
    if (rc != PSB_OK) return rc;
    g_dec[key] = d;
    *out = d;
    return PSB_OK;
}

template <int M>
int cols_go(const NufftParams& p, int nf, cudaStream_t s) {
    using C = NufftCfg<M>;
    static rt::PerDeviceOnce once;
    int rc0 = once.run([] {
        cudaError_t e = cudaFuncSetAttribute(nufft_cols_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem);
        if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("nufft columns: ") + cudaGetErrorString(e));
        return (int)PSB_OK;
    });
    if (rc0 != PSB_OK) return rc0;
    const long long n_tiles = (long long)(p.ny / C::W) * p.pair_count * nf;
    if (n_tiles > 0x7fffffffLL) return fail(PSB_ERR_UNSUPPORTED, "nufft columns: too many tiles per launch");
    const int sms = rt::sm_count();
    nufft_cols_kernel<M><<<dim3((unsigned)(n_tiles < sms ? n_tiles : sms)), C::kThreads, C::kSmem, s>>>(p, (int)n_tiles);
    ++launch_counter();
    return rt::check("nufft columns launch");
}

}  // namespace

void sf_mode_set(int mode) { g_sf_mode.store(mode < 0 ? 0 : (mode > 2 ? 2 : mode)); }
int sf_mode() { return g_sf_mode.load(); }

bool sf_nufft_supported(int ntypes, int nx, int ny) {
    if (nx != 512 && nx != 1024) return false;
    const int W = nx == 1024 ? 4 : 8;
    return ny % W == 0 && ntypes * (2 * nx / kBinCells) <= kMaxKeys;
}

// dense enough for the transform to beat the direct sum: atoms per slice pair and type (measured crossover)
bool sf_nufft_wanted(int ntypes, int nx, int ny, int n_atoms, int nz) {
    const int mode = g_sf_mode.load();
    if (mode == 1 || !sf_nufft_supported(ntypes, nx, ny)) return false;
    if (mode == 2) return true;
    const double per_pair_type = 2.0 * (double)n_atoms / (double)(nz > 0 ? nz : 1) / (double)ntypes;
    return nx >= 1024 && per_pair_type >= 350.0;      // measured at 1024 x 1024: equal cost near 320 atoms per pair and type
}

void sf_nufft_release() {
    std::lock_guard<std::mutex> lk(g_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : g_ws) {
        cudaSetDevice(kv.first.first);
        rt::dev_free(kv.second.block);
    }
    for (auto& kv : g_dec) {
        cudaSetDevice(kv.first.first);
        rt::dev_free(kv.second);
    }
    cudaSetDevice(cur);
    g_ws.clear();
    g_dec.clear();
}

namespace {
// workspace + tables of a call, resolved once per (device, owner stream); K1 fills it for the whole call
int nufft_params(const int* offsets, const unsigned int* ux, const unsigned int* uy, int cap, int nz, int ntypes, int nx, int ny,
                 int n_frames, const float* ff, cudaStream_t owner, bool may_grow, NufftParams* out) {
    if (!sf_nufft_supported(ntypes, nx, ny)) return fail(PSB_ERR_UNSUPPORTED, "nufft structure factor: unsupported grid");
    const int M = 2 * nx;
    NufftParams p;
    std::memset(&p, 0, sizeof(p));
    p.offsets = offsets; p.ux = ux; p.uy = uy; p.cap = cap; p.nz = nz; p.ntypes = ntypes; p.nx = nx; p.ny = ny;
    p.npairs = (nz + 1) / 2; p.nb = M / kBinCells; p.log_m = M == 2048 ? 11 : 10;
    p.ff = ff;
    int rc = dec_table(nx, &p.dec, owner);
    if (rc != PSB_OK) return rc;
    FftTables tb;
    int N = 0;
    bool blue = false;
    rc = get_fft_tables(M, &tb, &N, &blue, owner);
    if (rc != PSB_OK) return rc;
    if (blue || N != M) return fail(PSB_ERR_UNSUPPORTED, "nufft structure factor: power-of-two fine grid expected");
    p.tw = tb.tw;
    const int nkeys = ntypes * p.nb;
    const size_t n_xoff = (size_t)n_frames * p.npairs * (nkeys + 1), n_rec = (size_t)n_frames * cap;
    const size_t n_corner = (size_t)n_frames * p.npairs * ntypes * 2;
    const size_t need = n_xoff * sizeof(int) + 3 * n_rec * sizeof(unsigned int) + n_corner * sizeof(float) + n_rec * kTaps * sizeof(float) + 512;
    std::lock_guard<std::mutex> lk(g_mu);
    NufftWorkspace& w = g_ws[std::make_pair(rt::device(), owner)];
    if (need > w.bytes) {
        if (!may_grow) return fail(PSB_ERR_INVALID, "launch_sf_nufft without sf_nufft_prepare");
        cudaError_t e = cudaStreamSynchronize(owner);          // kernels of earlier calls may still read the old block
        if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("nufft workspace sync: ") + cudaGetErrorString(e));
        graph_cache_release();                                 // recorded launch sequences point at the block freed here
        rt::dev_free(w.block);
        w.block = rt::dev_alloc(need);
        w.bytes = w.block ? need : 0;
        if (!w.block) return PSB_ERR_NOMEM;
    }
    p.xoff = reinterpret_cast<int*>(w.block);
    p.rx = reinterpret_cast<unsigned int*>(p.xoff + n_xoff);
    p.ry = p.rx + n_rec;
    p.rpar = p.ry + n_rec;
    p.corner = reinterpret_cast<float*>(p.rpar + n_rec);
    // tap weights: 32 bytes per record, 16-byte aligned
    uintptr_t wt = reinterpret_cast<uintptr_t>(p.corner + n_corner);
    wt = (wt + 15) & ~(uintptr_t)15;
    p.rwt = reinterpret_cast<float*>(wt);
    *out = p;
    return PSB_OK;
}
}  // namespace

// once per psb_build_* call: order the atoms of every (frame, slice pair) by (type, x bin) and take the corner sums
int sf_nufft_prepare(const int* offsets, const unsigned int* ux, const unsigned int* uy, int cap, int nz, int ntypes, int nx,
                     int ny, int n_frames, const float* ff, cudaStream_t s, cudaStream_t owner) {
    NufftParams p;
    int rc = nufft_params(offsets, ux, uy, cap, nz, ntypes, nx, ny, n_frames, ff, owner, true, &p);
    if (rc != PSB_OK) return rc;
    const int nkeys = ntypes * p.nb;
    size_t prep_smem = (size_t)(nkeys + 1 + 256) * sizeof(int) + 4096 * sizeof(unsigned short) + 16;
    if (prep_smem < 256 * sizeof(double)) prep_smem = 256 * sizeof(double);
    static rt::PerDeviceOnce once;
    rc = once.run([] {
        cudaError_t e = cudaFuncSetAttribute(nufft_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("nufft prep: ") + cudaGetErrorString(e));
        return (int)PSB_OK;
    });
    if (rc != PSB_OK) return rc;
    nufft_prep_kernel<<<dim3(p.npairs, n_frames), 256, prep_smem, s>>>(p);
    ++launch_counter();
    return rt::check("nufft prep launch");
}

// per chunk: frames [frame0, frame0 + nf) x pairs [pair_begin, pair_begin + pair_count) -> out (nf, pair_count, nx, ny)
int launch_sf_nufft(const int* offsets, const unsigned int* ux, const unsigned int* uy, int cap, int nz, int ntypes, int nx,
                    int ny, int n_frames, int frame0, int nf, int pair_begin, int pair_count, const float* ff, float2* out,
                    cudaStream_t s, cudaStream_t owner) {
    NufftParams p;
    int rc = nufft_params(offsets, ux, uy, cap, nz, ntypes, nx, ny, n_frames, ff, owner, false, &p);
    if (rc != PSB_OK) return rc;
    p.frame0 = frame0; p.pair_begin = pair_begin; p.pair_count = pair_count; p.out = out;
    return 2 * nx == 2048 ? cols_go<2048>(p, nf, s) : cols_go<1024>(p, nf, s);
}

}  // namespace psb
