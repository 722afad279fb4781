// Structure factor of the projected potential fused with the inverse column transform
// (reference: src/multislice/potentials.py:319-337 -- the einsum over exp(-2 pi i kx x) exp(-2 pi i ky y) of the
// atoms of a slice, times the form factor, and the first half of ifft2).
//
// Same arithmetic as sf_fast.cu / StructureFactorPaired (quarter spectrum: four real sums per non-negative
// frequency slot, two slices packed per complex image, Nyquist lines riding in slot 0), but the work unit is
// shaped for the transform that follows instead of for the sum alone:
//
//   unit = (slice-pair image, group of GY ky slots)  ->  ALL kx slots x GY ky slots (1024 slots, 4 per thread)
//
// The mirror expansion of such a unit is exactly W = 2*GY complete spectrum columns (ky and ny-ky; the group
// holding ky = 0 also owns the Nyquist column), i.e. one tile of the column pass.  So the unit's sums are
// expanded into a shared-memory tile, inverse-transformed along kx in place (fast_fft.cuh) and stored as
// columns of the (x, ky) image: the spectrum never exists in global memory, the separate column pass
// (fast_cols_kernel<.., C_INVERSE>, one L2 round trip of the whole chunk) disappears, and because the unit is
// 16x finer than a (tile, pair) of sf_fast.cu the persistent grid balances for any chunk size.
//
// Persistent CTAs (2 per SM).  Thread 0 streams the phase-table rows of the unit's atoms global -> shared with
// cp.async.bulk through an S-stage full/empty mbarrier ring that runs ahead ACROSS unit boundaries (blocks of CH
// atom entries of the unit's contiguous entry range, regardless of segment boundaries): warps hand a stage back
// with one mbarrier arrive each, so there is no CTA barrier in the accumulation loop and only warp 0 (the producer)
// ever waits for the slowest consumer -- and it refills a stage D blocks after leaving it, when that wait is over.
// The segment offsets of all units of a CTA are staged in shared memory once, so no global-memory latency sits on
// the per-unit path.
#include "async_ptx.cuh"
#include "fast_fft.cuh"
#include "fast_path.h"
#include "pdl.cuh"
#include "potential_kernels.cuh"
#include "psb_rt.h"
#include "tables.h"

#include <cstring>
#include <mutex>
#include <string>

namespace psb {

namespace {

using fast::cpx;
using namespace aptx;

constexpr int kMaxUnitsPerCta = 64;

// tuning knobs (tools/microbench_potential.py times alternative builds through PSB_VARIANT_LIB)
#ifndef PSB_SFC_CH256
#define PSB_SFC_CH256 8
#endif
#ifndef PSB_SFC_STAGES
#define PSB_SFC_STAGES 4
#endif
#ifndef PSB_SFC_MINBLOCKS
#define PSB_SFC_MINBLOCKS 3
#endif
#ifndef PSB_SFC_DEFER
#define PSB_SFC_DEFER 1
#endif

template <int N>
struct SfcCfg {
    static constexpr int T = N / 16;                       // threads per column transform
    static constexpr int W = 256 / T;                      // columns per unit: 16 (N = 256), 8 (N = 512)
    static constexpr int GY = W / 2;                       // ky slots per unit
    static constexpr int NSX = N / 2;                      // kx slots (Nyquist rides in slot 0)
    static constexpr int XG = 256 / GY;                    // thread owns kx slots xg + XG*i, i < 4
    static constexpr int CH = (N == 256) ? PSB_SFC_CH256 : PSB_SFC_CH256 / 2;   // atom entries per ring stage
    static constexpr int S = PSB_SFC_STAGES;               // ring stages
    static constexpr int D = PSB_SFC_DEFER;                // a stage is refilled D blocks after warp 0 left it
    static_assert(D >= 0 && D < S, "refill deferral must leave at least one block in flight");
    static constexpr int kStage = CH * (NSX + GY);         // cpx per stage: [CH][NSX] x rows, then [CH][GY] y rows
    static constexpr int kPadRows = (W == 8) ? N / 16 : 0;
    static constexpr int kZ = (N + kPadRows) * W;          // spectrum tile = exchange buffer of the transform
    static_assert(4 * XG == NSX, "four kx slots per thread");
    static size_t smem(int ntypes) {
        return (size_t)(kZ + S * kStage) * sizeof(float2) + 2 * S * sizeof(uint64_t) +
               (size_t)kMaxUnitsPerCta * (2 * ntypes + 1) * sizeof(int);
    }
};

struct SfColsParams {
    const int* offsets;         // (nf, nseg+1), first frame of the chunk
    const unsigned int* ux;     // (nf, cap)
    const unsigned int* uy;
    int cap, nz, ntypes, ny;
    int pair_begin, pair_count;
    int n_units;                // nf * pair_count * (ny/2 / GY)
    float2* tabx;               // [nf][cap][NSX]      (cos, sin)(2 pi g u); slot 0 of the axis: (1, cos(pi n u))
    float2* taby;               // [nf][QG][cap][GY]
    float* snx;                 // [nf][cap]  Nyquist sine (corner term)
    float* sny;
    const float4* ff4;          // (ntypes, NSX, NSY) gathered form factors (sf_fast.cu: ff4_kernel)
    const float2* tw;           // staged twiddles of Plan<N, 16>
    float2* out;                // (nf * pair_count, N, ny): inverse transform along the first axis done
};

// (cos, sin)(2 pi g u) through the SFU; u is a 32-bit turn fraction so g*u wraps exactly and the argument handed
// to sin.approx / cos.approx is in [-pi, pi), where their absolute error is <= 2^-21.2.  Slot 0 of an even axis
// carries the Nyquist cosine in its (unused) sine component, its sine goes to *sn_out.
__device__ __forceinline__ float2 slot_phase_sfu(int g, unsigned int u, int n, float* sn_out) {
    const int ph = (int)((unsigned int)g * u);
    float s, c;
    __sincosf((float)ph * 1.4629180792671596e-9f, &s, &c);            // pi * 2^-31
    if (g == 0) {
        const int pq = (int)((unsigned int)(n / 2) * u);
        float s2, c2;
        __sincosf((float)pq * 1.4629180792671596e-9f, &s2, &c2);
        *sn_out = s2;
        return make_float2(1.f, c2);
    }
    return make_float2(c, s);
}

// K1: one warp per atom entry of the chunk
template <int N>
__global__ void __launch_bounds__(256) sfc_tables_kernel(const SfColsParams p) {
    using C = SfcCfg<N>;
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
    const int nsy = p.ny / 2, qg = nsy / C::GY;
    float sn = 0.f;
    float2* rowx = p.tabx + ((long long)fl * p.cap + e) * C::NSX;
    for (int g = lane; g < C::NSX; g += 32) rowx[g] = slot_phase_sfu(g, ux, N, &sn);
    if (lane == 0) p.snx[(long long)fl * p.cap + e] = sn;
    for (int g = lane; g < nsy; g += 32) {
        const float2 z = slot_phase_sfu(g, uy, p.ny, &sn);
        p.taby[(((long long)fl * qg + g / C::GY) * p.cap + e) * C::GY + g % C::GY] = z;
    }
    if (lane == 0) p.sny[(long long)fl * p.cap + e] = sn;
}

// exchange policy of the in-tile column transform: one buffer (the spectrum tile itself), CTA barriers
template <int N>
struct SfcXchg {
    cpx* z;
    int c;
    __device__ __forceinline__ cpx* buf(int) const { return z; }
    __device__ __forceinline__ int at(int q) const { return (SfcCfg<N>::W == 8 ? q + (q >> 4) : q) * SfcCfg<N>::W + c; }
    __device__ __forceinline__ void after_store(int) const { __syncthreads(); }
    __device__ __forceinline__ void after_load(int) const { __syncthreads(); }
};

// K2
template <int N>
__global__ void __launch_bounds__(256, PSB_SFC_MINBLOCKS) sf_cols_kernel(const SfColsParams p) {
    using C = SfcCfg<N>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* ztile = reinterpret_cast<cpx*>(smem_raw);                          // [kZ]
    cpx* ring = ztile + C::kZ;                                              // [S][kStage]
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + C::S * C::kStage);  // [S]
    uint64_t* empty = full + C::S;                                          // [S]
    int* uoff = reinterpret_cast<int*>(empty + C::S);                       // [units of this CTA][2*ntypes+1]

    const int tid = threadIdx.x;
    const int nt = p.ntypes, ustride = 2 * nt + 1;
    const int NY = p.ny, nsy = NY / 2, QG = nsy / C::GY;
    const int nseg = p.nz * nt;
    const int G = (int)gridDim.x;
    const int my_units = (p.n_units - (int)blockIdx.x + G - 1) / G;         // units u = blockIdx.x + ul*G

    pdl_trigger();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < C::S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 8);          // one arrive per warp
        }
        mbar_init_fence();
    }
    pdl_wait();          // offsets / tables come from the previous kernels of the chain

    // segment offsets of every unit of this CTA: uoff[ul][k] = off[2*m*nt + k], k = 0 .. 2*nt (clamped past the
    // last slice, which makes the missing second slice of an odd stack a run of empty segments)
    for (int w = tid; w < my_units * ustride; w += 256) {
        const int ul = w / ustride, k = w - ul * ustride;
        const int img = ((int)blockIdx.x + ul * G) / QG;
        const int fl = img / p.pair_count, m = p.pair_begin + img - fl * p.pair_count;
        int idx = 2 * m * nt + k;
        if (idx > nseg) idx = nseg;
        uoff[w] = p.offsets[(long long)fl * (nseg + 1) + idx];
    }
    __syncthreads();

    // ---- producer: warp 0 tracks the cursor (warp-uniform), its lane 0 issues ------------------------------
    const int warp = tid >> 5, lane = tid & 31;
    int p_ul = 0, p_pos = 0, p_seq = 0;
    auto issue = [&]() {
        int e0 = 0, e1 = 0;
        while (p_ul < my_units) {
            e0 = uoff[p_ul * ustride];
            e1 = uoff[p_ul * ustride + 2 * nt];
            if (e0 + p_pos < e1) break;
            ++p_ul;
            p_pos = 0;
        }
        if (p_ul >= my_units) return;
        const int e = e0 + p_pos;
        const int nc = e1 - e < C::CH ? e1 - e : C::CH;
        if (lane == 0) {
            const int u = (int)blockIdx.x + p_ul * G;
            const int img = u / QG, q = u - img * QG, fl = img / p.pair_count;
            const int st = p_seq % C::S;
            if (p_seq >= C::S) mbar_wait(&empty[st], (uint32_t)((p_seq / C::S - 1) & 1));   // every warp left its last use
            cpx* dst = ring + st * C::kStage;
            mbar_expect_tx(&full[st], (uint32_t)(nc * (C::NSX + C::GY) * sizeof(float2)));
            bulk_g2s(dst, p.tabx + ((long long)fl * p.cap + e) * C::NSX, (uint32_t)(nc * C::NSX * sizeof(float2)), &full[st]);
            bulk_g2s(dst + C::CH * C::NSX, p.taby + (((long long)fl * QG + q) * p.cap + e) * C::GY,
                     (uint32_t)(nc * C::GY * sizeof(float2)), &full[st]);
        }
        ++p_seq;
        p_pos += nc;
    };
    if (warp == 0) {
#pragma unroll 1
        for (int s = 0; s < C::S; ++s) issue();
    }

    // ---- consumer ------------------------------------------------------------------------------------------
    const int gyi = tid % C::GY, xg = tid / C::GY;      // accumulation: ky slot GY*q + gyi, kx slots xg + XG*i
    const int fc = tid % C::W, fj = tid / C::W;         // transform: column fc of the tile, line position fj
    const SfcXchg<N> xc{ztile, fc};
    auto zat = [](int kx) { return (C::W == 8 ? kx + (kx >> 4) : kx) * C::W; };
    int c_seq = 0;

#pragma unroll 1
    for (int ul = 0; ul < my_units; ++ul) {
        const int u = (int)blockIdx.x + ul * G;
        const int img = u / QG, q = u - img * QG, fl = img / p.pair_count;
        const int* uo = uoff + ul * ustride;
        const int E0 = uo[0], E1 = uo[2 * nt];
        const int gy = C::GY * q + gyi;
        const bool corner_warp = q == 0 && tid < 32;       // warp 0 of the unit holding (kx, ky) = (0, 0)
        const float* snx = p.snx + (long long)fl * p.cap;
        const float* sny = p.sny + (long long)fl * p.cap;

        float tot[4][4];        // [i][cc, ss, cs, sc], multiplied by the form factor
        float first[4][4];      // the pair's first slice while the second one accumulates
        int pos = E0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) tot[i][k] = 0.f;
#pragma unroll 1
            for (int t = 0; t < nt; ++t) {
                const int e = uo[h * nt + t + 1];
                if (uo[h * nt + t] == e) continue;
                float4 f[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) f[i] = __ldg(&p.ff4[((long long)t * C::NSX + xg + C::XG * i) * nsy + gy]);
                cpx P[4], Q[4];      // P = (cc, cs), Q = (sc, ss)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    P[i] = fast::c_make(0.f, 0.f);
                    Q[i] = fast::c_make(0.f, 0.f);
                }
                float corr = 0.f;
#pragma unroll 1
                while (pos < e) {
                    const int a0 = (pos - E0) & (C::CH - 1);
                    const int block_end = pos - a0 + C::CH;
                    const int run_end = e < block_end ? e : block_end;
                    const int a1 = a0 + run_end - pos;
                    const cpx* ex = ring + (c_seq % C::S) * C::kStage;
                    const cpx* ey = ex + C::CH * C::NSX + gyi;
                    ex += xg;
                    mbar_wait(&full[c_seq % C::S], (uint32_t)((c_seq / C::S) & 1));
#pragma unroll 4
                    for (int a = a0; a < a1; ++a) {
                        const cpx y = ey[a * C::GY];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const cpx x = ex[a * C::NSX + C::XG * i];
                            const cpx xcc = fast::c_make(fast::c_re(x), fast::c_re(x));
                            const cpx xss = fast::c_make(fast::c_im(x), fast::c_im(x));
                            P[i] = fast::fma2(xcc, y, P[i]);
                            Q[i] = fast::fma2(xss, y, Q[i]);
                        }
                    }
                    if (corner_warp)
                        for (int a = pos + tid; a < run_end; a += 32) corr += __ldg(&snx[a]) * __ldg(&sny[a]);
                    pos = run_end;
                    if (pos == block_end || pos == E1) {       // block consumed: hand the stage back
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty[c_seq % C::S]);
                        ++c_seq;
                        if (warp == 0 && c_seq > C::D) issue();
                    }
                }
                if (corner_warp) {
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) corr += __shfl_xor_sync(0xffffffffu, corr, d);
                    if (tid != 0) corr = 0.f;              // thread 0 owns slot (0, 0) of the corner unit as i = 0
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float cc = fast::c_re(P[i]), cs = fast::c_im(P[i]);
                    const float sc = fast::c_re(Q[i]);
                    float ss = fast::c_im(Q[i]);
                    if (i == 0) ss -= corr;
                    tot[i][0] += cc * f[i].x;
                    tot[i][1] += ss * f[i].y;
                    tot[i][2] += cs * f[i].z;
                    tot[i][3] += sc * f[i].w;
                }
            }
            if (h == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) first[i][k] = tot[i][k];
            }
        }

        // tot = second slice (B), first = first slice (A):  Z = S'_A + i*S'_B at the four mirror positions of a slot.
        // Tile columns: gyi <-> ky = gy, GY + gyi <-> ky = ny - gy (gy = 0: the Nyquist column ny/2).
        auto emit = [&](int kx, int col, float ar, float ai, float br, float bi) {
            ztile[zat(kx) + col] = fast::c_make(ar - bi, ai + br);
        };
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = xg + C::XG * i;
            const float acc_ = first[i][0], ass = first[i][1], acs = first[i][2], asc = first[i][3];
            const float bcc = tot[i][0], bss = tot[i][1], bcs = tot[i][2], bsc = tot[i][3];
            const int xm = x > 0 ? N - x : N / 2;
            const int cp = gyi, cm = C::GY + gyi;
            if (x > 0 && gy > 0) {
                emit(x, cp, acc_ - ass, -(acs + asc), bcc - bss, -(bcs + bsc));
                emit(xm, cp, acc_ + ass, -(acs - asc), bcc + bss, -(bcs - bsc));
                emit(x, cm, acc_ + ass, acs - asc, bcc + bss, bcs - bsc);
                emit(xm, cm, acc_ - ass, acs + asc, bcc - bss, bcs + bsc);
            } else if (x == 0 && gy > 0) {
                emit(0, cp, acc_, -acs, bcc, -bcs);
                emit(0, cm, acc_, acs, bcc, bcs);
                emit(xm, cp, asc, -ass, bsc, -bss);
                emit(xm, cm, asc, ass, bsc, bss);
            } else if (x > 0 && gy == 0) {
                emit(x, cp, acc_, -asc, bcc, -bsc);
                emit(xm, cp, acc_, asc, bcc, bsc);
                emit(x, cm, acs, -ass, bcs, -bss);
                emit(xm, cm, acs, ass, bcs, bss);
            } else {
                emit(0, cp, acc_, 0.f, bcc, 0.f);
                emit(0, cm, acs, 0.f, bcs, 0.f);
                emit(xm, cp, asc, 0.f, bsc, 0.f);
                emit(xm, cm, ass, 0.f, bss, 0.f);
            }
        }
        __syncthreads();

        // inverse transform of the tile's columns along kx, in place through the tile, straight to the image
        fast::Twiddles<N> tw;
        tw.load(p.tw, fj);
        cpx v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = ztile[xc.at(fj + e * C::T)];
        __syncthreads();
        const int g = C::GY * q + (fc < C::GY ? fc : fc - C::GY);
        const int ky = fc < C::GY ? g : (g == 0 ? NY / 2 : NY - g);
        cpx* dst = reinterpret_cast<cpx*>(p.out) + (long long)img * N * NY + (long long)fj * NY + ky;
        fast::line_fft<N, +1>([&](int e) { return v[e]; }, [&](int e, cpx a) { dst[(long long)e * C::T * NY] = a; }, tw, fj, xc, 0);
    }
}

// grow-only workspace for the phase tables, one per process (one process per GPU)
std::mutex g_mu;
void* g_ws = nullptr;
size_t g_ws_bytes = 0;

template <int N>
int sf_cols_go(SfColsParams p, int nf, const float2* tw, cudaStream_t s) {
    using C = SfcCfg<N>;
    const int nsy = p.ny / 2, qg = nsy / C::GY;
    const size_t nx_elems = (size_t)nf * p.cap * C::NSX, ny_elems = (size_t)nf * p.cap * nsy;
    const size_t sn_elems = (size_t)nf * p.cap;
    const size_t need = (nx_elems + ny_elems) * sizeof(float2) + 2 * sn_elems * sizeof(float) + 256;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (need > g_ws_bytes) {
            cudaError_t e = cudaStreamSynchronize(s);      // kernels of earlier chunks may still read the old block
            if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("sf workspace sync: ") + cudaGetErrorString(e));
            rt::dev_free(g_ws);
            g_ws = rt::dev_alloc(need);
            g_ws_bytes = g_ws ? need : 0;
            if (!g_ws) return PSB_ERR_NOMEM;
        }
        p.tabx = reinterpret_cast<float2*>(g_ws);
        p.taby = p.tabx + nx_elems;
        p.snx = reinterpret_cast<float*>(p.taby + ny_elems);
        p.sny = p.snx + sn_elems;
    }
    p.tw = tw;
    const size_t smem = C::smem(p.ntypes);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(sf_cols_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("sf+columns kernel: ") + cudaGetErrorString(e));
        smem_set = smem;
    }
    if (p.cap > 0) {
        cudaError_t e1 = pdl_launch(sfc_tables_kernel<N>, dim3((p.cap + 7) / 8, nf), dim3(256), 0, s, p);
        ++launch_counter();
        if (e1 != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("phase tables launch: ") + cudaGetErrorString(e1));
    }
    // the persistent kernel stages the offsets of at most kMaxUnitsPerCta units per CTA (sf_cols_supported checks;
    // chunks are a few hundred images at most: psb_build_transmission sizes them to stay L2-resident)
    const int slots = 2 * rt::sm_count();
    const int imgs_total = nf * p.pair_count;
    if ((long long)imgs_total * qg > (long long)kMaxUnitsPerCta * slots)
        return fail(PSB_ERR_UNSUPPORTED, "sf+columns kernel: chunk too large for one launch");
    p.n_units = imgs_total * qg;
    const int grid = p.n_units < slots ? p.n_units : slots;
    cudaError_t e = pdl_launch(sf_cols_kernel<N>, dim3(grid), dim3(256), smem, s, p);
    ++launch_counter();
    if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("sf+columns launch: ") + cudaGetErrorString(e));
    return PSB_OK;
}

}  // namespace

void sf_cols_release() {
    std::lock_guard<std::mutex> lk(g_mu);
    rt::dev_free(g_ws);
    g_ws = nullptr;
    g_ws_bytes = 0;
}

bool sf_cols_supported(int nx, int ny, int n_img) {
    if (!fast_slice_supported(nx, ny) || fast_path_level() < 2) return false;
    const int qg = (ny / 2) / (nx == 256 ? 8 : 4);
    return (long long)n_img * qg <= (long long)kMaxUnitsPerCta * 2 * rt::sm_count();
}

int launch_sf_cols(const int* offsets, const unsigned int* ux, const unsigned int* uy, int cap, int nz, int ntypes, int nx,
                   int ny, int pair_begin, int pair_count, int nf, const float4* ff4, float2* out, cudaStream_t s) {
    SfColsParams p;
    std::memset(&p, 0, sizeof(p));
    p.offsets = offsets; p.ux = ux; p.uy = uy; p.cap = cap; p.nz = nz; p.ntypes = ntypes; p.ny = ny;
    p.pair_begin = pair_begin; p.pair_count = pair_count; p.ff4 = ff4; p.out = out;
    FftTables tb;
    int N = 0;
    bool blue = false;
    int rc = get_fft_tables(nx, &tb, &N, &blue, s);
    if (rc != PSB_OK) return rc;
    if (blue || N != nx) return fail(PSB_ERR_UNSUPPORTED, "sf+columns kernel needs a power-of-two column");
    if (nx == 256) return sf_cols_go<256>(p, nf, tb.tw, s);
    if (nx == 512) return sf_cols_go<512>(p, nf, tb.tw, s);
    return fail(PSB_ERR_UNSUPPORTED, "sf+columns kernel: unsupported grid");
}

}  // namespace psb
