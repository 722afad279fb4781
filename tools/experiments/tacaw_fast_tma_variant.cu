// EXPERIMENT, not part of libpsb: tacaw_fast.cu as it stood with two measured-and-rejected additions --
//   * tacaw_tma_kernel: persistent CTAs fed by 3-D tensor copies, two tiles in flight per CTA (PSB_TACAW_TMA=1).
//     C3 quarter 4.94 against 4.90 TB/s for the plain kernel with the same radix-10 stages, C2 2.4 against 3.0,
//     T = 2000 1.46 against 2.24 (profiles/r2z_tacaw_history.txt): bytes in flight were not what held the kernel back,
//     the balance of its stages was (320 threads for 640 butterfly-pixels per stage: 5.61 TB/s);
//   * a single-buffer form of the same kernel for whole-SM tiles (T = 2000): 1.375 against 1.374 ms -- load and transform of a
//     128 KB tile stay serialised either way, and shared memory has no room for a second one;
//   * L2 prefetch of the next whole-SM tile (PSB_TACAW_AHEAD): T = 2000 1.72 against 1.40 ms.
// TACAW time-axis transform for frame counts of the form 2^a 3^b 5^c (every BASELINE.json configuration: 20, 100, 500,
// 2000) -- reference: src/postprocessing/tacaw_data.py:61-106,
//     I[p, w, k] = | fftshift_t FFT_t( psi[p, t, k] - <psi[p, ., k]>_t ) |^2 .
//
// The generic pass (line_pass.cuh, TW) treats every pixel's time series as a strided line (element stride = one
// image) and transforms non-powers of two by Bluestein: three padded power-of-two transforms per line.  At
// configuration scale (C3: 6.7e9 elements) that is compute-bound at ~15x the 12 B/element this row needs.  Here
//
//   * a CTA owns PX adjacent pixels x all T frames: loads are T segments of PX*8 contiguous bytes, stores T segments
//     of PX*4 bytes (whole 32 B sectors for PX >= 8), the tile sits in shared memory as [t][px] in between;
//   * the transform is an in-place mixed-radix decimation-in-frequency (radices 5, 4, 3, 2 from a host factorisation):
//     stage s splits blocks of B elements into R sub-blocks, y[k1] = DFT_R(x)[k1] * W_B^(n' k1), no ping-pong buffer,
//     one CTA barrier per stage; the digit-reversed order is undone by a permutation table that also carries the
//     fftshift, so the output sweep is a plain coalesced store of |.|^2;
//   * the first stage reads its butterfly inputs straight from global memory and the last stage writes |.|^2 straight
//     back, so a T = R1 R2 R3 series crosses shared memory four times in all;
//   * psi - <psi>_t differs from psi in the zero-frequency bin only, where it is exactly 0: no mean pass, bin 0 is
//     written as 0 (the static part rides along the k = 0 chain of the decimation only; its cancellation error in the
//     other bins, <= 6e-8 * |static| * T / R_last, is of the order of the input's own float32 quantisation).
//
// Twiddles exp(-2 pi i n / T) and the permutation are built once per T in float64 on the host and cached on the device.
#include "pdl.cuh"
#include "psb_rt.h"
#include "tacaw_fast.h"
#include "tacaw_stages.cuh"

#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace psb {

namespace {

using namespace tw;

struct TwFastParams {
    const float2* wf;
    long long stride_probe, stride_frame;   // elements
    int T;
    long long npix;
    float* out;                             // (P, T, npix)
    const float2* tw;                       // [T]  exp(-2 pi i n / T)
    const int* perm;                        // [T]  storage position -> shifted frequency index
    int nfac;
    int fac[kMaxFactors];
    int ahead;                              // whole-SM tiles: tile x + ahead is pulled into L2 while tile x is transformed (0: off)
};

// kThreads: 256 when several tiles fit one SM's shared memory, 1024 when a tile (long series) has the SM to itself
template <int PX, int kThreads, int BIG>
__global__ void __launch_bounds__(kThreads) tacaw_fast_kernel(const TwFastParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* data = reinterpret_cast<float2*>(smem_raw);             // [T][PX]
    const int T = p.T;
    const int px = threadIdx.x % PX;
    const long long gpx = (long long)blockIdx.x * PX + px;
    const bool live = gpx < p.npix;
    const float2* src = p.wf + (long long)blockIdx.y * p.stride_probe + gpx;
    float* dst = p.out + (long long)blockIdx.y * T * p.npix + gpx;

    if (p.nfac == 1) {
        PSB_TW_DISPATCH(p.fac[0], (last_stage<R, PX, kThreads, true>(threadIdx.x, data, src, p.stride_frame, live, p.perm, dst, p.npix, T)));
        return;
    }
    PSB_TW_DISPATCH(p.fac[0], (first_stage<R, PX, kThreads>(threadIdx.x, data, src, p.stride_frame, live, p.tw, T)));
    if constexpr (kThreads >= 512) {
        // A tile that has the SM to itself leaves HBM idle while its middle stages run.  Pull the rows of the tile that
        // the next CTA on this SM will most likely take into L2 meanwhile (whichever SM ends up with it, L2 is shared).
        const long long nx = (long long)blockIdx.x + p.ahead;
        if (p.ahead > 0 && nx < gridDim.x) {
            const float2* nsrc = p.wf + (long long)blockIdx.y * p.stride_probe + nx * PX;
            for (int t = threadIdx.x; t < T; t += kThreads)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc + (long long)t * p.stride_frame));
        }
    }
    __syncthreads();
    int B = T / p.fac[0];
    for (int s = 1; s < p.nfac - 1; ++s) {
        PSB_TW_DISPATCH(p.fac[s], (mid_stage<R, PX, kThreads>(threadIdx.x, data, p.tw, T, B)));
        B /= p.fac[s];
        __syncthreads();
    }
    PSB_TW_DISPATCH(p.fac[p.nfac - 1], (last_stage<R, PX, kThreads, false>(threadIdx.x, data, src, p.stride_frame, live, p.perm, dst, p.npix, T)));
}

// ---- persistent variant: tiles arrive through the async proxy, two in flight per CTA ---------------------------------
// The kernel above loads a tile with ordinary loads in its first stage and then computes with nothing in flight; with four
// CTAs per SM in different phases an SM keeps ~30 KB of reads outstanding, two thirds of what HBM's latency-bandwidth
// product asks for (ncu r2v: 0.61 of the copy bandwidth, 3.8 long-scoreboard stalls per issue).  Here a CTA owns two tile
// buffers: while it transforms one in place, the tensor copy (3-D map: pixel, frame, probe) of the tile after the next one
// lands in the other, so a whole tile per CTA is always in flight regardless of what the warps are doing.  Out-of-range
// pixels of the last tile are zero-filled by the copy; every stage is the in-place middle stage (the first with B = T).
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
__device__ __forceinline__ void tensor3d_g2s_once(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

struct TwTmaParams {
    int T;
    long long npix;
    float* out;                             // (P, T, npix)
    const float2* tw;
    const int* perm;
    int nfac;
    int fac[kMaxFactors];
    int box_rows, n_boxes;                  // T = box_rows * n_boxes, box_rows <= 256 (tensor-copy box limit)
    int buf_elems;                          // float2 per tile buffer (T * PX rounded up to 128 bytes)
    long long tiles_per_probe, n_tiles;
};

template <int PX, int kThreads, int BIG>
__global__ void __launch_bounds__(kThreads) tacaw_tma_kernel(const __grid_constant__ CUtensorMap map, const TwTmaParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* bufs = reinterpret_cast<float2*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(bufs + 2 * (size_t)p.buf_elems);
    const int T = p.T;
    const unsigned tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint64_t stream_once;                   // the wave functions are read exactly once
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream_once));
    auto issue = [&](long long tile, int b) {
        const int probe = (int)(tile / p.tiles_per_probe);
        const int c0 = (int)(tile - probe * p.tiles_per_probe) * PX;
        mbar_expect_tx(&full[b], (uint32_t)((size_t)T * PX * sizeof(float2)));
        for (int i = 0; i < p.n_boxes; ++i)
            tensor3d_g2s_once(bufs + (size_t)b * p.buf_elems + (size_t)i * p.box_rows * PX, &map, c0, i * p.box_rows, probe, &full[b], stream_once);
    };
    const long long G = gridDim.x;
    long long tile = blockIdx.x;
    if (tid == 0) {
        if (tile < p.n_tiles) issue(tile, 0);
        if (tile + G < p.n_tiles) issue(tile + G, 1);
    }
    for (uint32_t it = 0; tile < p.n_tiles; tile += G, ++it) {
        const int b = (int)(it & 1u);
        float2* data = bufs + (size_t)b * p.buf_elems;
        mbar_wait(&full[b], (it >> 1) & 1u);
        int B = T;
        for (int s = 0; s < p.nfac - 1; ++s) {
            PSB_TW_DISPATCH(p.fac[s], (mid_stage<R, PX, kThreads>(tid, data, p.tw, T, B)));
            B /= p.fac[s];
            __syncthreads();
        }
        const long long probe = tile / p.tiles_per_probe;
        const long long gpx = (tile - probe * p.tiles_per_probe) * PX + tid % PX;
        float* dst = p.out + probe * T * p.npix + gpx;
        PSB_TW_DISPATCH(p.fac[p.nfac - 1], (last_stage<R, PX, kThreads, false>(tid, data, nullptr, 0, gpx < p.npix, p.perm, dst, p.npix, T)));
        // the buffer goes back to the async proxy: order this thread's generic-proxy accesses before the copy that follows the barrier
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0 && tile + 2 * G < p.n_tiles) issue(tile + 2 * G, b);
    }
}

// ---- host side -------------------------------------------------------------------------------------------
struct TwTables {
    float2* tw = nullptr;
    int* perm = nullptr;
    int nfac = 0;
    int fac[kMaxFactors] = {0};
};

std::mutex g_mu;
std::map<std::pair<int, int>, TwTables> g_tables;      // (device, T)

int get_tables(int T, TwTables* out, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_mu);
    const auto key = std::make_pair(rt::device(), T);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) {
        *out = it->second;
        return PSB_OK;
    }
    TwTables tb;
    if (!factorise(T, tb.fac, &tb.nfac)) return fail(PSB_ERR_UNSUPPORTED, "tacaw fast path: frame count is not 2^a 3^b 5^c");
    std::vector<float2> tw(T);
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < T; ++n) {
        const double a = -two_pi * (double)n / (double)T;
        tw[n] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    std::vector<int> perm(T);
    build_perm(T, tb.fac, tb.nfac, perm.data());
    tb.tw = static_cast<float2*>(rt::dev_alloc(T * sizeof(float2)));
    tb.perm = static_cast<int*>(rt::dev_alloc(T * sizeof(int)));
    if (!tb.tw || !tb.perm) return PSB_ERR_NOMEM;
    int rc = rt::h2d(tb.tw, tw.data(), T * sizeof(float2), s);
    if (rc == PSB_OK) rc = rt::h2d(tb.perm, perm.data(), T * sizeof(int), s);
    if (rc != PSB_OK) return rc;
    g_tables[key] = tb;
    *out = tb;
    return PSB_OK;
}

inline int radix_class(const int* fac, int nfac) {      // 0: radices up to 5, 1: up to 10, 2: up to 20
    int c = 0;
    for (int i = 0; i < nfac; ++i) c = fac[i] > 10 ? 2 : (fac[i] > 5 && c < 1 ? 1 : c);
    return c;
}

template <int PX, int kThreads, int BIG>
int go_as(const TwFastParams& p, int n_probes, cudaStream_t s) {
    const size_t smem = (size_t)p.T * PX * sizeof(float2);
    static std::atomic<size_t> smem_set[64];          // per device ordinal (the attribute is per device), zero-initialised
    const int d = rt::device() & 63;
    if (smem > smem_set[d].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(tacaw_fast_kernel<PX, kThreads, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("tacaw fast path: ") + cudaGetErrorString(e));
        smem_set[d].store(smem, std::memory_order_release);
    }
    const long long tiles = (p.npix + PX - 1) / PX;
    tacaw_fast_kernel<PX, kThreads, BIG><<<dim3((unsigned)tiles, (unsigned)n_probes), kThreads, smem, s>>>(p);
    ++launch_counter();
    return rt::check("tacaw fast launch");
}

template <int PX, int kThreads>
int go(const TwFastParams& p, int n_probes, cudaStream_t s) {
    // radix-10 / 20 butterflies want up to 127 registers: a whole-SM tile runs them with 512 threads instead of 1024
    constexpr int kBigThreads = kThreads > 512 ? 512 : kThreads;
    switch (radix_class(p.fac, p.nfac)) {
        case 2: return go_as<PX, kBigThreads, 2>(p, n_probes, s);
        case 1: return go_as<PX, kBigThreads, 1>(p, n_probes, s);
        default: return go_as<PX, kThreads, 0>(p, n_probes, s);
    }
}


// ---- persistent tensor-copy variant: planning and launch -------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#ifndef PSB_TACAW_320
#define PSB_TACAW_320 1
#endif
#ifndef PSB_TACAW_TMA_THREADS
#define PSB_TACAW_TMA_THREADS 512
#endif
constexpr int kTmaThreads2 = PSB_TACAW_TMA_THREADS;          // two CTAs per SM
constexpr int kTmaThreads1 = 2 * PSB_TACAW_TMA_THREADS;      // one CTA per SM

struct TmaPlan {
    int px = 0, threads = 0, ctas_per_sm = 0, box_rows = 0, n_boxes = 0, buf_elems = 0;
    size_t smem = 0;
};

// tile width and CTA shape of the persistent variant: the widest tile of which two buffers fit twice per SM, else 8 (or 4)
// pixels with the SM to one CTA.  px == 0: not applicable (single-stage transform, tile too large, unaligned input).
TmaPlan tma_plan(const float2* wf, long long stride_probe, long long stride_frame, int T, int nfac) {
    TmaPlan pl;
    static const bool enabled = [] {
        const char* e = std::getenv("PSB_TACAW_TMA");
        return e && e[0] == '1';
    }();
    if (!enabled || nfac < 2) return pl;
    if ((reinterpret_cast<uintptr_t>(wf) & 15) || (stride_frame & 1) || (stride_probe & 1)) return pl;     // 16-byte tensor-map strides
    // boxes of a tile: at most 256 frames each (tensor-copy limit), equal, landing on 128-byte boundaries
    auto boxes = [&](int px) {
        for (int nb = (T + 255) / 256; nb <= 16; ++nb)
            if (T % nb == 0 && ((T / nb) * px) % 16 == 0) return nb;
        return 0;
    };
    auto fits = [&](int px, size_t budget) {
        const size_t buf = (((size_t)T * px * sizeof(float2)) + 127) / 128 * 128;
        return boxes(px) > 0 && 2 * buf + 64 <= budget;
    };
    for (int px : {64, 32, 16, 8})
        if (fits(px, 110u << 10)) { pl.px = px; pl.threads = kTmaThreads2; pl.ctas_per_sm = 2; break; }
    if (!pl.px)
        for (int px : {8, 4})
            if (fits(px, 220u << 10)) { pl.px = px; pl.threads = kTmaThreads1; pl.ctas_per_sm = 1; break; }
    if (!pl.px) return pl;
    pl.n_boxes = boxes(pl.px);
    pl.box_rows = T / pl.n_boxes;
    pl.buf_elems = (int)(((((size_t)T * pl.px * sizeof(float2)) + 127) / 128 * 128) / sizeof(float2));
    pl.smem = 2 * (size_t)pl.buf_elems * sizeof(float2) + 2 * sizeof(uint64_t);
    return pl;
}

int encode_wf_map(CUtensorMap* map, const float2* wf, long long stride_probe, long long stride_frame, int n_probes, int T, long long npix,
                  int box_px, int box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !ptr)
            return fail(PSB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    // the wave functions as a 3-D tensor of 8-byte elements: dim0 = pixel, dim1 = frame, dim2 = probe
    const cuuint64_t gdim[3] = {(cuuint64_t)npix, (cuuint64_t)T, (cuuint64_t)n_probes};
    const cuuint64_t gstride[2] = {(cuuint64_t)stride_frame * sizeof(float2), (cuuint64_t)(n_probes > 1 ? stride_probe : stride_frame * T) * sizeof(float2)};
    const cuuint32_t box[3] = {(cuuint32_t)box_px, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<float2*>(wf), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return PSB_ERR_UNSUPPORTED;          // odd strides / sizes: the caller falls back to the first kernel
    return PSB_OK;
}

template <int PX, int kThreads, int BIG>
int go_tma_as(const CUtensorMap& map, const TwTmaParams& p, const TmaPlan& pl, cudaStream_t s) {
    static std::atomic<size_t> smem_set[64];
    const int d = rt::device() & 63;
    if (pl.smem > smem_set[d].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(tacaw_tma_kernel<PX, kThreads, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("tacaw tensor-copy path: ") + cudaGetErrorString(e));
        smem_set[d].store(pl.smem, std::memory_order_release);
    }
    const long long slots = (long long)pl.ctas_per_sm * rt::sm_count();
    const unsigned grid = (unsigned)(p.n_tiles < slots ? p.n_tiles : slots);
    tacaw_tma_kernel<PX, kThreads, BIG><<<grid, kThreads, pl.smem, s>>>(map, p);
    ++launch_counter();
    return rt::check("tacaw tensor-copy launch");
}

template <int PX, int kThreads>
int go_tma(const CUtensorMap& map, const TwTmaParams& p, const TmaPlan& pl, cudaStream_t s) {
    constexpr int kBigThreads = kThreads > 512 ? 512 : kThreads;
    switch (radix_class(p.fac, p.nfac)) {
        case 2: return go_tma_as<PX, kBigThreads, 2>(map, p, pl, s);
        case 1: return go_tma_as<PX, kBigThreads, 1>(map, p, pl, s);
        default: return go_tma_as<PX, kThreads, 0>(map, p, pl, s);
    }
}

}  // namespace

bool tacaw_fast_supported(int T) {
    int fac[kMaxFactors], n = 0;
    return T >= 2 && factorise(T, fac, &n) && pick_px(T) > 0;
}

void tacaw_fast_release() {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_tables) {
        rt::dev_free(kv.second.tw);
        rt::dev_free(kv.second.perm);
    }
    g_tables.clear();
}

int launch_tacaw_fast(const float2* wf, long long stride_probe, long long stride_frame, int n_probes, int n_frames,
                      long long npix, float* intensity, cudaStream_t s) {
    TwTables tb;
    int rc = get_tables(n_frames, &tb, s);
    if (rc != PSB_OK) return rc;
    TwFastParams p;
    std::memset(&p, 0, sizeof(p));
    p.wf = wf; p.stride_probe = stride_probe; p.stride_frame = stride_frame; p.T = n_frames; p.npix = npix;
    p.out = intensity; p.tw = tb.tw; p.perm = tb.perm; p.nfac = tb.nfac;
    std::memcpy(p.fac, tb.fac, sizeof(p.fac));
    static const int ahead_waves = [] {
        const char* e = std::getenv("PSB_TACAW_AHEAD");
        return e ? std::atoi(e) : 1;
    }();
    p.ahead = ahead_waves * rt::sm_count();
    if (n_probes == 0 || npix == 0) return PSB_OK;
    const TmaPlan pl = tma_plan(wf, stride_probe, stride_frame, n_frames, tb.nfac);
    if (pl.px) {
        CUtensorMap map;
        if (encode_wf_map(&map, wf, stride_probe, stride_frame, n_probes, n_frames, npix, pl.px, pl.box_rows) == PSB_OK) {
            TwTmaParams q;
            std::memset(&q, 0, sizeof(q));
            q.T = n_frames; q.npix = npix; q.out = intensity; q.tw = tb.tw; q.perm = tb.perm; q.nfac = tb.nfac;
            std::memcpy(q.fac, tb.fac, sizeof(q.fac));
            q.box_rows = pl.box_rows; q.n_boxes = pl.n_boxes; q.buf_elems = pl.buf_elems;
            q.tiles_per_probe = (npix + pl.px - 1) / pl.px;
            q.n_tiles = q.tiles_per_probe * n_probes;
            switch (pl.px) {
                case 64: return go_tma<64, kTmaThreads2>(map, q, pl, s);
                case 32: return go_tma<32, kTmaThreads2>(map, q, pl, s);
                case 16: return go_tma<16, kTmaThreads2>(map, q, pl, s);
                case 8: return pl.ctas_per_sm == 2 ? go_tma<8, kTmaThreads2>(map, q, pl, s) : go_tma<8, kTmaThreads1>(map, q, pl, s);
                default: return go_tma<4, kTmaThreads1>(map, q, pl, s);
            }
        }
    }
    const bool one_per_sm = whole_sm(n_frames);      // 32 warps on a tile that has the SM to itself
    switch (pick_px(n_frames)) {
        case 64: {
            // radix-10 stages of T = 100 k: T/10 butterflies x 64 pixels split evenly over 320 threads but not over 256
            const long long items = (long long)(n_frames / 10) * 64;
            if (tb.fac[0] == 10 && items % 320 == 0 && items % 256 != 0 && PSB_TACAW_320) return go_as<64, 320, 1>(p, n_probes, s);
            return go<64, 256>(p, n_probes, s);
        }
        case 32: return go<32, 256>(p, n_probes, s);
        case 16: return go<16, 256>(p, n_probes, s);
        case 8: return one_per_sm ? go<8, 1024>(p, n_probes, s) : go<8, 256>(p, n_probes, s);
        case 4: return go<4, 1024>(p, n_probes, s);
        default: return fail(PSB_ERR_UNSUPPORTED, "tacaw fast path: frame count too large for a shared-memory tile");
    }
}

}  // namespace psb
