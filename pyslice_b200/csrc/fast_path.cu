// Fused slice-step kernels for 256-, 512- and 1024-point lines: the steady state of psb_propagate
// (reference: src/multislice/multislice.py:281-294, one loop iteration = one row pass + one column pass).
//
//   row pass   psi[x, ky] -> FFT_y( t_s[x, y] * IFFT_y( psi[x, ky] ) )                  (in place)
//   col pass   psi[x, ky] -> IFFT_x( Px[kx] * FFT_x( Py[ky] * psi[x, ky] ) )            (in place)
//
// with the Fresnel propagator split as P[kx, ky] = Px[kx] * Py[ky]; Py commutes with the x transforms and ky is
// the column a thread of the column pass owns, so it is one register per tile there (applied on load).
// Both are persistent kernels sized to the SM count:
//
//   * inputs arrive through the async proxy: cp.async.bulk (TMA, 1-D) global -> shared, completion on an
//     mbarrier; the copy for the NEXT tile is issued as soon as the current tile's landing buffer has been
//     read into registers, so L2/HBM latency hides behind the two line FFTs of the current tile;
//   * row pass: a line (2 KB / 4 KB) belongs to the 16 / 32 lanes of ONE warp, so every warp is its own
//     pipeline (private landing + exchange buffers, private mbarriers, __syncwarp only; no CTA barrier in
//     the loop); 16 warps per SM.  A 1024-point line (8 KB) belongs to a PAIR of warps that synchronise on
//     their own named barrier (bar.sync id, 64), eight such pipelines per SM;
//   * column pass: a tile is 16 (8 for N = 512 and 1024) adjacent columns so global segments are 128 B (64 B);
//     256 threads (512 for N = 1024), one CTA barrier per FFT stage exchange thanks to two alternating
//     exchange buffers; 2 CTAs per SM (1 for N = 1024);
//   * arithmetic is packed fp32x2 with per-thread persistent twiddles (fast_fft.cuh); the propagator
//     factors sit in shared memory for the whole kernel.
//
// Algorithmic traffic per slice step and image: psi is read and written once per pass (L2-resident for the
// batch sizes engine.py picks), t is read once from HBM by the row pass.
#include "fast_fft.cuh"
#include "fast_path.h"
#include "pdl.cuh"
#include "psb_rt.h"
#include "tables.h"

#include <cuda.h>

#include <atomic>
#include <cstring>
#include <string>

namespace psb {

namespace {

std::atomic<int> g_fast_enabled{1};

using fast::cpx;

// ---- async-proxy plumbing (PTX) --------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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

// same, with an L2 eviction-priority hint: the transmission stack streams through L2 once (evict first) while
// psi must stay resident between the passes
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// 2-D tiled tensor copy (TMA): box of the tensor map at element coordinates (c0 = column, c1 = row)
__device__ __forceinline__ void tensor2d_g2s(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// exp(i*x) for the two halves of a packed pair at once (the transmission epilogue: two slices share one pixel of
// the paired inverse transform).  Packed Cody-Waite reduction by whole turns (k = rint(x / 2 pi) through the
// magic-number add, r = x - k*2pi in two FMAs), then the SFU: sin.approx / cos.approx on r in [-pi, pi], where
// their absolute error is bounded by 2^-21.2 (4.2e-7).  The SFU pipe is otherwise idle in this kernel; the
// degree-7/8 polynomial + quadrant selection this replaces cost ~60 FMA/ALU-pipe instructions per pixel pair and
// made the epilogue more expensive than the transform (ncu r1h: issue-bound at 62 %, 33 us against a 22 us HBM
// write floor).  Error budget: uniform noise of that half-width injected into every t of a 512-slice stack moves
// the exit wave by 7.8e-6 rel-L2 and the TACAW cube by 1.0e-5 (tolerances 1e-4 / 1e-3; /tmp experiment recorded in
// DESIGN.md section 4.2).
__device__ __forceinline__ void pair_cis(cpx x, cpx& ea, cpx& eb) {
    const cpx magic = fast::c_make(12582912.f, 12582912.f);                         // 1.5 * 2^23
    cpx k = fast::fma2(x, fast::c_make(0.15915494309189535f, 0.15915494309189535f), magic);
    k = fast::sub2(k, magic);
    cpx r = fast::fma2(k, fast::c_make(-6.2831854820251465f, -6.2831854820251465f), x);
    r = fast::fma2(k, fast::c_make(1.7484555e-7f, 1.7484555e-7f), r);               // 2 pi = 6.28318548... - 1.7484555e-7
    float sa, ca, sb, cb;
    __sincosf(fast::c_re(r), &sa, &ca);
    __sincosf(fast::c_im(r), &sb, &cb);
    ea = fast::c_make(ca, sa);
    eb = fast::c_make(cb, sb);
}

// exp(i*x) for one phase (R_STEP_PHASE): the scalar form of pair_cis
__device__ __forceinline__ cpx cis1(float x) {
    float k = fmaf(x, 0.15915494309189535f, 12582912.f) - 12582912.f;
    float r = fmaf(k, -6.2831854820251465f, x);
    r = fmaf(k, 1.7484555e-7f, r);
    float sn, cs;
    __sincosf(r, &sn, &cs);
    return fast::c_make(cs, sn);
}

// ---- row pass ------------------------------------------------------------------------------------------
struct RowPassParams {
    float2* psi;                 // (n_img, nx, N) contiguous, transformed in place
    const float2* t;             // transmission slice of frame 0 (already offset to slice z)
    long long t_frame_stride;    // elements between consecutive frames of t
    int probes;                  // images per frame (image = frame*probes + probe)
    int nx;                      // lines per image
    const float2* tw;            // staged twiddle table of Plan<N, 16>
    long long n_units;           // n_img * nx / lines-per-warp
    // R_TRANSMIT: image = frame*pair_count + ml holds V_{2m} + i*V_{2m+1}, m = pair_begin + ml; the transmission
    // function of slice s of that frame goes to t_out[(frame*pair_nz + s)*nx*N + ...] (v_out likewise, optional)
    float2* t_out;
    float* v_out;
    float scale, sigma;
    int pair_count, pair_nz, pair_begin;
};

// R_PHASE / R_STEP_PHASE: the transmission stack as float32 phases sigma*V (4 B per pixel and slice instead of 8):
// R_PHASE writes them (no sincos), R_STEP_PHASE reads them and evaluates exp(i*phase) on the SFU as it multiplies.
// Worth it when few probes share a frame's stack (plane-wave runs): the stack is written once and read once per probe.
enum RowMode { R_STEP = 0, R_TRANSMIT = 1, R_PHASE = 2, R_STEP_PHASE = 3 };
constexpr bool row_mode_steps(int mode) { return mode == R_STEP || mode == R_STEP_PHASE; }

// transmission landing buffers per warp (R_STEP): with 2 the copy of unit n+2's rows is issued while unit n is
// transformed, a full unit of lead more against the HBM round trip, at the price of 12 instead of 16 warps per SM
#ifndef PSB_ROW_TBUFS
#define PSB_ROW_TBUFS 1
#endif
// the same for the phase-format slice step: its rows are half the size, so two buffers still leave room for 16 warps
#ifndef PSB_ROW_TBUFS_PHASE
#define PSB_ROW_TBUFS_PHASE 1
#endif

// warps per CTA of the phase-format slice step (R_STEP_PHASE): its transmission landing buffer is half the size, which
// leaves room for 20 warps (5 per scheduler) in shared memory and, at <= 96 registers, in the register file
#ifndef PSB_ROW_WARPS_PHASE
#define PSB_ROW_WARPS_PHASE 16
#endif

// phases of the slice (R_STEP_PHASE) loaded straight into registers with streaming loads at the start of a unit (1) instead
// of TMA -> landing buffer -> LDS (0): 8 B per pixel less through the L1 / shared-memory pipe, 16 more live registers
// exp(i*phase) of the phase-format slice step evaluated for two pixels at a time (1) or one (0)
#ifndef PSB_ROW_PAIR_CIS
#define PSB_ROW_PAIR_CIS 1
#endif
#ifndef PSB_ROW_PHASE_LDG
#define PSB_ROW_PHASE_LDG 0
#endif

template <int N, int MODE = 0>
struct RowCfg {
    static constexpr int T = N / 16;               // threads per line
    static constexpr int kGroupThreads = T > 32 ? T : 32;      // threads that share buffers and synchronise: a warp, or a warp pair (N = 1024)
    static constexpr int LPW = kGroupThreads / T;  // lines per group (unit of work)
    static constexpr int NP = N + N / 16;          // padded exchange pitch
    static constexpr int kTBufs = (MODE == 0) ? PSB_ROW_TBUFS : ((MODE == 3) ? PSB_ROW_TBUFS_PHASE : 1);
    static constexpr int kWarps = (MODE == 3) ? PSB_ROW_WARPS_PHASE : ((kTBufs == 1) ? 16 : 12);
    static constexpr int kThreads = 32 * kWarps;
    static constexpr int kGroups = kThreads / kGroupThreads;
    static constexpr int kLand = LPW * N;          // float2 per landing buffer (4 KB; 8 KB for N = 1024)
    static constexpr int kTLand = (MODE == 3) ? kLand / 2 : kLand;     // transmission rows of a unit: float phases or complex t
    static constexpr int kX = LPW * NP;
    static constexpr int kWarpElems = kLand + kTBufs * kTLand + kX;    // per group
    static constexpr int kBars = 1 + kTBufs;
    static constexpr size_t kSmem = (size_t)kGroups * kWarpElems * sizeof(float2) + kGroups * kBars * sizeof(uint64_t);
    static constexpr uint32_t kBytes = kLand * sizeof(float2);
    static constexpr uint32_t kTBytes = kLand * (MODE == 3 ? sizeof(float) : sizeof(float2));     // transmission rows of a unit
    static_assert(kThreads % kGroupThreads == 0 && (kGroupThreads == 32 || kGroups <= 15), "one named barrier per warp pair");
};

// barrier over the threads of one group: the warp itself, or the named barrier of a warp pair
template <int THREADS>
__device__ __forceinline__ void group_sync(int group) {
    if constexpr (THREADS == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(THREADS) : "memory");
}

template <int N>
struct RowXchg {
    cpx* x;
    int c;
    int group;
    __device__ __forceinline__ cpx* buf(int) const { return x; }
    __device__ __forceinline__ int at(int q) const { return c * RowCfg<N>::NP + q + (q >> 4); }
    __device__ __forceinline__ void after_store(int) const { group_sync<RowCfg<N>::kGroupThreads>(group); }
    __device__ __forceinline__ void after_load(int) const { group_sync<RowCfg<N>::kGroupThreads>(group); }
    __device__ __forceinline__ void mid_sync(int) const { group_sync<RowCfg<N>::kGroupThreads>(group); }
};

template <int N, int MODE>
__global__ void __launch_bounds__(RowCfg<N, MODE>::kThreads, 1) fast_rows_kernel(const RowPassParams p) {
    using C = RowCfg<N, MODE>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* sm = reinterpret_cast<cpx*>(smem_raw);
    // group = the warp (or warp pair) that owns a unit; provably warp-uniform: bulk copies take uniform operands
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x / C::kGroupThreads), 0);
    const int lane = threadIdx.x % C::kGroupThreads;
    cpx* land_psi = sm + (size_t)warp * C::kWarpElems;
    cpx* land_t = land_psi + C::kLand;                       // [kTBufs][kLand]
    cpx* xb = land_t + C::kTBufs * C::kTLand;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)C::kGroups * C::kWarpElems) + C::kBars * warp;
    uint64_t* mb_psi = bars;
    uint64_t* mb_t = bars + 1;                               // [kTBufs]

    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < C::kBars; ++b) mbar_init(bars + b, 1);
        mbar_init_fence();
    }
    pdl_trigger();
    const int c = lane / C::T, j = lane % C::T;
    fast::Twiddles<N> tw;
    tw.load(p.tw, j);
    const RowXchg<N> xc{xb, c, warp};
    __syncthreads();
    pdl_wait();          // everything above reads constant tables only; the images come from the previous kernel

    const uint64_t stream_once = l2_policy_evict_first();
    // unit / image arithmetic in 32 bits with shifts: nx is 256 or 512 here, n_units < 2^31 (64-bit division is
    // a ~100-instruction subroutine, and it sat in every prefetch)
    const int GW = (int)gridDim.x * C::kGroups;
    int u = (int)blockIdx.x * C::kGroups + warp;
    const int n_units = (int)p.n_units;
    const int upi_shift = 31 - __clz(p.nx / C::LPW);          // log2(units per image)
    const int upi_mask = (1 << upi_shift) - 1;

    auto t_rows = [&](int unit) -> const void* {
        const unsigned img = (unsigned)unit >> upi_shift;
        const int row0 = (unit & upi_mask) * C::LPW;
        const long long o = (long long)(img / (unsigned)p.probes) * p.t_frame_stride + row0 * N;
        if (MODE == R_STEP_PHASE) return reinterpret_cast<const float*>(p.t) + o;
        return p.t + o;
    };
    auto issue = [&](int unit, bool want_psi, bool want_t, int tb = 0) {
        if (want_psi) {
            mbar_expect_tx(mb_psi, C::kBytes);
            if (MODE == R_TRANSMIT || MODE == R_PHASE)      // last use of the chunk's rows: do not let them displace the next chunk's
                bulk_g2s_hint(land_psi, p.psi + (long long)unit * C::kLand, C::kBytes, mb_psi, stream_once);
            else
                bulk_g2s(land_psi, p.psi + (long long)unit * C::kLand, C::kBytes, mb_psi);
        }
        if (row_mode_steps(MODE) && want_t && !(MODE == R_STEP_PHASE && PSB_ROW_PHASE_LDG)) {
            mbar_expect_tx(mb_t + tb, C::kTBytes);
            bulk_g2s_hint(land_t + tb * C::kTLand, t_rows(unit), C::kTBytes, mb_t + tb, stream_once);
        }
    };
    if (u < n_units && lane == 0) {
        issue(u, true, true, 0);
#pragma unroll
        for (int b = 1; b < C::kTBufs; ++b)
            if (u + b * GW < n_units) issue(u + b * GW, false, true, b);
    }

    for (uint32_t it = 0; u < n_units; u += GW, ++it) {
        const uint32_t parity = it & 1u;
        const int un = u + GW;
        const bool next = un < n_units;
        const int tb = (int)(it % C::kTBufs);                        // this unit's transmission buffer
        const uint32_t t_parity = (it / C::kTBufs) & 1u;
        const int ut = u + C::kTBufs * GW;                           // the unit whose rows refill it
        mbar_wait(mb_psi, parity);
        // The landing buffers go back to the async proxy only from inside the transforms, after the first
        // exchange: its stores depend on every value loaded from them, so those loads have completed by then.
        // (Issuing right after a __syncwarp let the next unit's TMA overwrite words whose LDS was still in flight.)
        const cpx* lp = land_psi + c * N + j;
        const cpx* lt = land_t + tb * C::kTLand + c * N + j;
        if constexpr (row_mode_steps(MODE)) {
            cpx v[16];
            const float* ltf = reinterpret_cast<const float*>(land_t + tb * C::kTLand) + c * N + j;      // R_STEP_PHASE: float rows
            float tph[16];
            if constexpr (MODE == R_STEP_PHASE && PSB_ROW_PHASE_LDG) {                  // in flight during the inverse transform
                const float* gph = reinterpret_cast<const float*>(t_rows(u)) + c * N + j;
#pragma unroll
                for (int e = 0; e < 16; ++e) tph[e] = __ldcs(gph + e * C::T);
            }
            // psi[x, ky] -> IFFT_y -> * t[x, y]
            fast::line_fft<N, +1>(
                [&](int e) { return lp[e * C::T]; },
                [&](int e, cpx a) {
                    if constexpr (MODE == R_STEP_PHASE && PSB_ROW_PHASE_LDG) v[e] = fast::cmulp(a, cis1(tph[e]));
                    else if constexpr (MODE == R_STEP_PHASE && PSB_ROW_PAIR_CIS) {
                        // the last butterfly layer hands out positions k, k + 4, k + 8, k + 12: take them two at a time so the
                        // range reduction of exp(i*phase) runs on the packed pipe (4 instead of 8 fp32 instructions per pair)
                        if (((e >> 2) & 1) == 0) {
                            v[e] = a;
                        } else {
                            cpx ea, eb;
                            pair_cis(fast::c_make(ltf[(e - 4) * C::T], ltf[e * C::T]), ea, eb);
                            v[e - 4] = fast::cmulp(v[e - 4], ea);
                            v[e] = fast::cmulp(a, eb);
                        }
                    }
                    else if constexpr (MODE == R_STEP_PHASE) v[e] = fast::cmulp(a, cis1(ltf[e * C::T]));
                    else v[e] = fast::cmulp(a, lt[e * C::T]);
                },
                tw, j, xc, 0,
                [&]() {
                    if (lane == 0 && next) issue(un, true, false);
                },
                [&]() {
                    if constexpr (!(MODE == R_STEP_PHASE && PSB_ROW_PHASE_LDG)) mbar_wait(mb_t + tb, t_parity);
                });
            // -> FFT_y -> global
            cpx* dst = reinterpret_cast<cpx*>(p.psi) + (long long)u * C::kLand + c * N + j;
            fast::line_fft<N, -1>([&](int e) { return v[e]; }, [&](int e, cpx a) { dst[e * C::T] = a; }, tw, j, xc, 0, [&]() {
                if (lane == 0 && ut < n_units) issue(ut, false, true, tb);
            });
        } else {
            // potentials.py:336-342 + multislice.py:281-282: V = Re/Im(IFFT2) * scale, t = exp(i*sigma*V), two slices per image
            const unsigned img = (unsigned)u >> upi_shift;
            const int row = (u & upi_mask) * C::LPW + c;
            const unsigned fr = img / (unsigned)p.pair_count;
            const int m = p.pair_begin + (int)(img - fr * (unsigned)p.pair_count);
            const long long img_elems = (long long)p.nx * N;
            const long long o = ((long long)fr * p.pair_nz + 2 * m) * img_elems + row * N + j;
            const bool has_b = 2 * m + 1 < p.pair_nz;
            const cpx scale2 = fast::c_make(p.scale, p.scale), sigma2 = fast::c_make(p.sigma, p.sigma);
            cpx* ta_out = reinterpret_cast<cpx*>(p.t_out) + o;
            fast::line_fft<N, +1>(
                [&](int e) { return lp[e * C::T]; },
                [&](int e, cpx a) {
                    const cpx V2 = fast::mul2(a, scale2);         // (V_2m, V_2m+1)
                    if constexpr (MODE == R_PHASE) {              // phases sigma*V of the two slices, no sincos
                        const cpx ph = fast::mul2(V2, sigma2);
                        __stcs(p.v_out + o + e * C::T, fast::c_re(ph));
                        if (has_b) __stcs(p.v_out + o + img_elems + e * C::T, fast::c_im(ph));
                        return;
                    }
                    cpx ta, tb;
                    pair_cis(fast::mul2(V2, sigma2), ta, tb);
                    // streaming stores: t (134 MB per 64 MB chunk of spectra) must not evict the L2-resident chunk
                    __stcs(ta_out + e * C::T, ta);
                    if (p.v_out) __stcs(p.v_out + o + e * C::T, fast::c_re(V2));
                    if (has_b) {
                        __stcs(ta_out + img_elems + e * C::T, tb);
                        if (p.v_out) __stcs(p.v_out + o + img_elems + e * C::T, fast::c_im(V2));
                    }
                },
                tw, j, xc, 0,
                [&]() {
                    if (lane == 0 && next) issue(un, true, false);
                });
        }
    }
}

// ---- column pass ---------------------------------------------------------------------------------------
struct ColPassParams {
    float2* psi;                 // (n_img, N, NY) contiguous, transformed in place along the first image axis
    const float2* px;            // [N] propagator factor along the line (includes 1/(nx*ny))
    const float2* py;            // [NY] propagator factor of the column (C_PROPAGATE)
    const float2* tw;
    long long n_tiles;           // n_img * NY / W
};

// exchange buffers of the column pass: 2 alternate (one CTA barrier per exchange, 2 CTAs per SM by shared memory);
// 1 needs a second barrier per exchange but lets 3 CTAs share an SM
#ifndef PSB_COL_XBUFS
#define PSB_COL_XBUFS 2
#endif
constexpr int kColXBufs = PSB_COL_XBUFS;
// Px[kx] of the thread's 16 line positions in registers for the whole kernel (1) instead of 16 shared-memory loads per tile (0)
#ifndef PSB_COL_PX_REGS
#define PSB_COL_PX_REGS 1      // measured: +2.6 % at 256^2, +2 % at 512^2 (profiles/r2t_slice_step_variants.txt)
#endif

// 1024-point columns: 512 threads and 8-column tiles in one CTA per SM (default), or 256 threads and 4-column tiles in
// two CTAs per SM (-DPSB_COL1024_THREADS=256)
#ifndef PSB_COL1024_THREADS
#define PSB_COL1024_THREADS 512
#endif
// 512-point columns: 256 threads and 8-column tiles (64-byte global segments) in two CTAs per SM (default), or 512 threads
// and 16-column tiles (128-byte segments) in one CTA per SM (-DPSB_COL512_THREADS=512)
#ifndef PSB_COL512_THREADS
#define PSB_COL512_THREADS 256
#endif
// 256-point columns: 256 threads and 16-column tiles in two CTAs per SM (default), or 128 threads and 8-column tiles in four
#ifndef PSB_COL256_THREADS
#define PSB_COL256_THREADS 256
#endif
// 1024-point columns, alternative shape (1): two 256-thread CTAs per SM, each landing 8-column tiles (64-byte rows for the
// tensor copy) and transforming them as two halves of 4 columns through ONE exchange buffer -- two independent CTAs fill
// each other's barrier bubbles where the single 512-thread CTA idles (ncu r2v: issue slots 34 % busy)
#ifndef PSB_COL1024_HALVES
#define PSB_COL1024_HALVES 0
#endif
#ifndef PSB_COL1024_TABLES_SMEM
#define PSB_COL1024_TABLES_SMEM 1
#endif

// The inverse-only pass of the potential build does one transform per tile (ncu r2w: issue slots 29 % busy, 4.5
// long-scoreboard stalls per issue with 2 CTAs per SM).  1: one exchange buffer, 3 CTAs per SM -- measured equal
// (2.178 against 2.174 ms per 16 frames of C2, profiles/r2x_potential_chunk_sweep.txt), so the propagate pass's shape stays
#ifndef PSB_COL_INV_3CTAS
#define PSB_COL_INV_3CTAS 0
#endif

enum ColMode { C_PROPAGATE = 0, C_INVERSE = 1 };

template <int N, int MODE>
struct ColCfg {
    static constexpr int T = N / 16;
    static constexpr int kHalves = (N == 1024 && PSB_COL1024_HALVES) ? 2 : 1;      // column groups of a tile transformed one after the other
    static constexpr int kXBufs = (kHalves == 2 || (MODE == C_INVERSE && N < 1024 && PSB_COL_INV_3CTAS)) ? 1 : kColXBufs;
    static constexpr int kThreads = N == 1024 ? (kHalves == 2 ? 256 : PSB_COL1024_THREADS) : (N == 512 ? PSB_COL512_THREADS : PSB_COL256_THREADS);
    static constexpr int kCtasPerSm = N == 1024 ? (kThreads == 512 ? 1 : 2)
                                    : ((N == 512 && kThreads == 512) ? 1 : (kXBufs == 1 ? 3 : 2) * (N == 256 ? 256 / kThreads : 1));
    static constexpr int W = kThreads / T;                            // columns transformed together: 16 (N=256), 8 (N=512, 1024), 4 (1024, 256 threads)
    static constexpr int WL = W * kHalves;                            // columns per tile (landing buffer, tensor-copy box)
    static constexpr int kPadRows = (W < 16) ? N / 16 : 0;            // keeps narrow rows conflict-free
    static constexpr int kLand = N * WL;                              // float2 (32 KB; 64 KB for N = 1024)
    static constexpr int kX = (N + kPadRows) * W;
    // Px, Py staged per CTA.  1024 with one 512-thread CTA per SM: 64 + 2 x 68 + 16 KB = 216 KB still fits (ncu r2v: with the
    // tables read through L1 the kernel's top stall was long-scoreboard, 4.1 per issue); two 256-thread CTAs: through L1;
    // the two-halves shape: Px staged, Py (one value per thread and half) through L1
    static constexpr bool kTablesInSmem = MODE == C_PROPAGATE && (N < 1024 || kHalves == 2 || (PSB_COL1024_TABLES_SMEM && kThreads == 512));
    static constexpr bool kPyInSmem = kTablesInSmem && kHalves == 1;
    static constexpr size_t kSmem = (size_t)(kLand + kXBufs * kX + (kTablesInSmem ? N : 0)) * sizeof(float2) + 2 * sizeof(uint64_t);      // + NY*8 for Py, added at launch
    static constexpr int kBoxRows = 256;                              // TMA box limit per dimension
    static constexpr uint32_t kBytes = kLand * sizeof(float2);
    static constexpr size_t smem_for(int ny) { return kSmem + (kPyInSmem ? (size_t)ny * sizeof(float2) : 0); }
};

// Exchange policy of the column pass: two alternating buffers, one CTA barrier per exchange.  The first
// barrier of a tile doubles as the "landing buffer is free" point: thread 0 then issues the next tile's TMA.
template <int N, int MODE>
struct ColXchg {
    using C = ColCfg<N, MODE>;
    cpx* b0;
    cpx* b1;
    int c;
    const CUtensorMap* map;
    uint64_t* mb;
    cpx* land;
    int next_c0, next_r0;        // tensor coordinates of the next tile; next_c0 < 0: nothing to prefetch
    int hook_i;                  // index of the tile's first exchange
    __device__ __forceinline__ cpx* buf(int i) const { return (C::kXBufs == 2 && (i & 1)) ? b1 : b0; }
    __device__ __forceinline__ int at(int q) const {
        return (C::W < 16 ? q + (q >> 4) : q) * C::W + c;
    }
    __device__ __forceinline__ void after_store(int i) const {
        __syncthreads();
        if (i == hook_i && next_c0 >= 0 && threadIdx.x == 0) {      // first barrier of the tile
            mbar_expect_tx(mb, C::kBytes);
#pragma unroll
            for (int h = 0; h < N / C::kBoxRows; ++h)
                tensor2d_g2s(land + h * C::kBoxRows * C::WL, map, next_c0, next_r0 + h * C::kBoxRows, mb);
        }
    }
    __device__ __forceinline__ void after_load(int) const {
        if (C::kXBufs == 1) __syncthreads();
    }
    __device__ __forceinline__ void mid_sync(int) const { __syncthreads(); }
};

template <int N, int NY, int MODE>
__global__ void __launch_bounds__(ColCfg<N, MODE>::kThreads, ColCfg<N, MODE>::kCtasPerSm) fast_cols_kernel(const __grid_constant__ CUtensorMap tmap, const ColPassParams p) {
    using C = ColCfg<N, MODE>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* land = reinterpret_cast<cpx*>(smem_raw);
    cpx* xb0 = land + C::kLand;
    cpx* xb1 = xb0 + (C::kXBufs - 1) * C::kX;
    cpx* spx = xb1 + C::kX;
    cpx* spy = spx + (C::kTablesInSmem ? N : 0);
    uint64_t* mb = reinterpret_cast<uint64_t*>(spy + (C::kPyInSmem ? NY : 0));

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(mb, 1);
        mbar_init_fence();
    }
    pdl_trigger();
    if constexpr (MODE == C_PROPAGATE && C::kTablesInSmem) {
        for (int i = tid; i < N; i += C::kThreads) spx[i] = reinterpret_cast<const cpx*>(p.px)[i];
        if constexpr (C::kPyInSmem)
            for (int i = tid; i < NY; i += C::kThreads) spy[i] = reinterpret_cast<const cpx*>(p.py)[i];
    }
    const cpx* gpx = reinterpret_cast<const cpx*>(p.px);
    const cpx* gpy = reinterpret_cast<const cpx*>(p.py);
    cpx pxr[16];
    constexpr bool kPxRegs = PSB_COL_PX_REGS && N < 1024;      // 1024-point columns have no registers to spare
    if constexpr (MODE == C_PROPAGATE && kPxRegs) {
#pragma unroll
        for (int e = 0; e < 16; ++e) pxr[e] = gpx[tid / C::W + e * C::T];
    }

    const int c = tid % C::W, j = tid / C::W;
    fast::Twiddles<N> tw;
    tw.load(p.tw, j);
    __syncthreads();
    pdl_wait();          // constant tables above, images below
    constexpr int kTilesPerImg = NY / C::WL;

    long long tile = blockIdx.x;
    const long long G = gridDim.x;
    ColXchg<N, MODE> xc{xb0, xb1, c, &tmap, mb, land, -1, 0, 0};
    if (tile < p.n_tiles && tid == 0) {
        mbar_expect_tx(mb, C::kBytes);
#pragma unroll
        for (int h = 0; h < N / C::kBoxRows; ++h)
            tensor2d_g2s(land + h * C::kBoxRows * C::WL, &tmap, (int)(tile % kTilesPerImg) * C::WL,
                         (int)(tile / kTilesPerImg) * N + h * C::kBoxRows, mb);
    }
    for (uint32_t it = 0; tile < p.n_tiles; tile += G, ++it) {
        const long long nt = tile + G;
        const int next_c0 = nt < p.n_tiles ? (int)(nt % kTilesPerImg) * C::WL : -1;
        xc.next_r0 = nt < p.n_tiles ? (int)(nt / kTilesPerImg) * N : 0;
        mbar_wait(mb, it & 1u);
#pragma unroll 1
        for (int half = 0; half < C::kHalves; ++half) {
            // the landing buffer goes back to the tensor copy once its last columns have been read
            xc.next_c0 = half == C::kHalves - 1 ? next_c0 : -1;
            const int col0 = (int)(tile % kTilesPerImg) * C::WL + half * C::W;
            const cpx* lp = land + j * C::WL + half * C::W + c;
            cpx* dst = reinterpret_cast<cpx*>(p.psi) + (tile / kTilesPerImg) * ((long long)N * NY) + col0 + j * NY + c;
            if constexpr (MODE == C_PROPAGATE) {
                cpx v[16];
                // * Py[ky] -> FFT_x -> * Px[kx] -> IFFT_x
                xc.hook_i = 0;
                const cpx pyc = C::kPyInSmem ? spy[col0 + c] : __ldg(gpy + col0 + c);
                fast::line_fft<N, -1>([&](int e) { return fast::cmulp(lp[e * C::T * C::WL], pyc); },
                                      [&](int e, cpx a) {
                                          if constexpr (kPxRegs) v[e] = fast::cmulp(a, pxr[e]);
                                          else v[e] = fast::cmulp(a, C::kTablesInSmem ? spx[j + e * C::T] : __ldg(gpx + j + e * C::T));
                                      },
                                      tw, j, xc, 0);
                xc.next_c0 = -1;
                fast::line_fft<N, +1>([&](int e) { return v[e]; }, [&](int e, cpx a) { dst[e * C::T * NY] = a; }, tw, j, xc,
                                      fast::exchanges<N>());
            } else {
                // one transform per tile: alternate the exchange buffer between tiles so a single barrier per
                // exchange still orders the reuse (N = 512 already alternates inside the transform)
                xc.hook_i = fast::exchanges<N>() == 1 ? (int)(it & 1u) : 0;
                fast::line_fft<N, +1>([&](int e) { return lp[e * C::T * C::WL]; }, [&](int e, cpx a) { dst[e * C::T * NY] = a; }, tw, j,
                                      xc, xc.hook_i);
            }
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------------
template <class K>
int ensure_smem(K kernel, size_t bytes, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return PSB_OK;
}

int twiddles_for(int n, const float2** tw, cudaStream_t s) {
    FftTables tb;
    int N = 0;
    bool blue = false;
    int rc = get_fft_tables(n, &tb, &N, &blue, s);
    if (rc != PSB_OK) return rc;
    if (blue || N != n) return fail(PSB_ERR_UNSUPPORTED, "fast path needs a power-of-two line");
    *tw = tb.tw;
    return PSB_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not a link-time dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_cols_map(CUtensorMap* map, float2* psi, long long rows_total, int ny, int box_cols, int box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !ptr)
            return fail(PSB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    // psi as a 2-D tensor of 8-byte elements: dim0 = column (ny), dim1 = row over all images
    const cuuint64_t gdim[2] = {(cuuint64_t)ny, (cuuint64_t)rows_total};
    const cuuint64_t gstride[1] = {(cuuint64_t)ny * sizeof(float2)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, psi, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PSB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return PSB_OK;
}

template <int N, int MODE>
int rows_go(const RowPassParams& p, cudaStream_t s) {
    using C = RowCfg<N, MODE>;
    static rt::PerDeviceOnce once;
    int rc0 = once.run([] { return ensure_smem(fast_rows_kernel<N, MODE>, C::kSmem, "fast row pass"); });
    if (rc0 != PSB_OK) return rc0;
    long long want = (p.n_units + C::kGroups - 1) / C::kGroups;
    const int sms = rt::sm_count();
    const int grid = (int)(want < sms ? want : sms);
    cudaError_t e = pdl_launch(fast_rows_kernel<N, MODE>, dim3(grid), dim3(C::kThreads), C::kSmem, s, p);
    ++launch_counter();
    if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("fast row pass launch: ") + cudaGetErrorString(e));
    return PSB_OK;
}

template <int N, int NY, int MODE>
int cols_go(float2* psi, int n_img, const float2* px, const float2* py, const float2* tw, cudaStream_t s) {
    using C = ColCfg<N, MODE>;
    static rt::PerDeviceOnce once;
    int rc0 = once.run([] { return ensure_smem(fast_cols_kernel<N, NY, MODE>, C::smem_for(NY), "fast column pass"); });
    if (rc0 != PSB_OK) return rc0;
    // the tensor map depends on the buffer and the batch size only: keep the last one of this host thread (a map is
    // plain host data passed by value at launch, so a per-thread cache needs no lock and cannot cross devices: device
    // pointers of different GPUs never compare equal under unified addressing)
    static thread_local CUtensorMap map;
    static thread_local float2* map_psi = nullptr;
    static thread_local int map_img = -1;
    if (map_psi != psi || map_img != n_img) {
        int rc = encode_cols_map(&map, psi, (long long)n_img * N, NY, C::WL, C::kBoxRows);
        if (rc != PSB_OK) return rc;
        map_psi = psi;
        map_img = n_img;
    }
    ColPassParams p;
    p.psi = psi; p.px = px; p.py = py; p.tw = tw;
    p.n_tiles = (long long)n_img * (NY / C::WL);
    const long long slots = (long long)C::kCtasPerSm * rt::sm_count();
    const int grid = (int)(p.n_tiles < slots ? p.n_tiles : slots);
    cudaError_t e = pdl_launch(fast_cols_kernel<N, NY, MODE>, dim3(grid), dim3(C::kThreads), C::smem_for(NY), s, map, p);
    ++launch_counter();
    if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("fast column pass launch: ") + cudaGetErrorString(e));
    return PSB_OK;
}

}  // namespace

void fast_path_enable(int level) { g_fast_enabled.store(level <= 0 ? 0 : 1); }
bool fast_path_enabled() { return g_fast_enabled.load() != 0; }
int fast_path_level() { return g_fast_enabled.load(); }

static bool fast_line(int n) { return n == 256 || n == 512 || n == 1024; }

bool fast_slice_supported(int nx, int ny) {
    if (!g_fast_enabled.load()) return false;
    return fast_line(nx) && fast_line(ny);
}

// row passes: dispatch on the line length ny and fill in the unit count
template <int MODE>
static int rows_dispatch(RowPassParams& p, int n_img, int nx, int ny, cudaStream_t s) {
    int rc = twiddles_for(ny, &p.tw, s);
    if (rc != PSB_OK) return rc;
    p.nx = nx;
    if (ny == 256) {
        p.n_units = (long long)n_img * nx / RowCfg<256, MODE>::LPW;
        return rows_go<256, MODE>(p, s);
    }
    if (ny == 512) {
        p.n_units = (long long)n_img * nx / RowCfg<512, MODE>::LPW;
        return rows_go<512, MODE>(p, s);
    }
    if (ny == 1024) {
        p.n_units = (long long)n_img * nx / RowCfg<1024, MODE>::LPW;
        return rows_go<1024, MODE>(p, s);
    }
    return fail(PSB_ERR_UNSUPPORTED, "fast row pass: unsupported line length");
}

int launch_fast_rows(float2* psi, int n_img, int nx, int ny, const float2* t_slice, long long t_frame_stride,
                     int probes, cudaStream_t s) {
    RowPassParams p;
    std::memset(&p, 0, sizeof(p));
    p.psi = psi; p.t = t_slice; p.t_frame_stride = t_frame_stride; p.probes = probes;
    return rows_dispatch<R_STEP>(p, n_img, nx, ny, s);
}

int launch_fast_rows_phase(float2* psi, int n_img, int nx, int ny, const float* phase_slice, long long t_frame_stride,
                           int probes, cudaStream_t s) {
    RowPassParams p;
    std::memset(&p, 0, sizeof(p));
    p.psi = psi; p.t = reinterpret_cast<const float2*>(phase_slice); p.t_frame_stride = t_frame_stride; p.probes = probes;
    return rows_dispatch<R_STEP_PHASE>(p, n_img, nx, ny, s);
}

int launch_fast_rows_phase_out(float2* pairs, int n_img, int nx, int ny, float scale, float sigma, float* phase_out,
                               int pair_count, int pair_nz, int pair_begin, cudaStream_t s) {
    RowPassParams p;
    std::memset(&p, 0, sizeof(p));
    p.psi = pairs; p.probes = 1;
    p.v_out = phase_out; p.scale = scale; p.sigma = sigma;
    p.pair_count = pair_count; p.pair_nz = pair_nz; p.pair_begin = pair_begin;
    return rows_dispatch<R_PHASE>(p, n_img, nx, ny, s);
}

int launch_fast_rows_transmit(float2* pairs, int n_img, int nx, int ny, float scale, float sigma, float2* t_out,
                              float* v_out, int pair_count, int pair_nz, int pair_begin, cudaStream_t s) {
    RowPassParams p;
    std::memset(&p, 0, sizeof(p));
    p.psi = pairs; p.probes = 1;
    p.t_out = t_out; p.v_out = v_out; p.scale = scale; p.sigma = sigma;
    p.pair_count = pair_count; p.pair_nz = pair_nz; p.pair_begin = pair_begin;
    return rows_dispatch<R_TRANSMIT>(p, n_img, nx, ny, s);
}

template <int MODE, int N>
static int cols_dispatch_ny(float2* psi, int n_img, int ny, const float2* px, const float2* py, const float2* tw, cudaStream_t s) {
    if (ny == 256) return cols_go<N, 256, MODE>(psi, n_img, px, py, tw, s);
    if (ny == 512) return cols_go<N, 512, MODE>(psi, n_img, px, py, tw, s);
    if (ny == 1024) return cols_go<N, 1024, MODE>(psi, n_img, px, py, tw, s);
    return fail(PSB_ERR_UNSUPPORTED, "fast column pass: unsupported grid");
}

template <int MODE>
static int cols_dispatch(float2* psi, int n_img, int nx, int ny, const float2* px, const float2* py, cudaStream_t s) {
    const float2* tw = nullptr;
    int rc = twiddles_for(nx, &tw, s);
    if (rc != PSB_OK) return rc;
    if (nx == 256) return cols_dispatch_ny<MODE, 256>(psi, n_img, ny, px, py, tw, s);
    if (nx == 512) return cols_dispatch_ny<MODE, 512>(psi, n_img, ny, px, py, tw, s);
    if (nx == 1024) return cols_dispatch_ny<MODE, 1024>(psi, n_img, ny, px, py, tw, s);
    return fail(PSB_ERR_UNSUPPORTED, "fast column pass: unsupported grid");
}

int launch_fast_cols(float2* psi, int n_img, int nx, int ny, const float2* px, const float2* py, cudaStream_t s) {
    return cols_dispatch<C_PROPAGATE>(psi, n_img, nx, ny, px, py, s);
}

int launch_fast_cols_inverse(float2* imgs, int n_img, int nx, int ny, cudaStream_t s) {
    return cols_dispatch<C_INVERSE>(imgs, n_img, nx, ny, nullptr, nullptr, s);
}

}  // namespace psb
