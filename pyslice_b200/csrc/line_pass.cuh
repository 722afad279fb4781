// The line-pass kernel: every FFT-shaped operator of the hot path is one instantiation of
//     load line -> [mean-subtract] -> [FFT a] -> [pointwise multiply] -> [FFT b] -> store variant
// over a tile of W lines per CTA, lines held in registers in the strided layout of fft_core.cuh.
//
//   ROWS orientation: a line is contiguous in memory (the y / ky axis of an (nx, ny) image),
//   COLS orientation: a line is strided by the row length (the x / kx axis, or the frame axis of
//                     the TACAW time FFT); lanes run across W neighbouring lines so global
//                     accesses are still W*8-byte contiguous segments.
//
// Instantiations used (reference file:line of the torch call(s) each one fuses):
//   R1  rows  psi = FFT_y(t0 * probe)                       multislice.py:281-286 + first half of :292
//   R   rows  psi = FFT_y(t_s * IFFT_y(psi))                second half of :294, :281-286, first half of :292
//   C   cols  psi = IFFT_x(P * FFT_x(psi))                  second half of :292, :293, first half of :294
//   CX  cols  out = fftshift(FFT_x(psi))                    calculators.py:286-287 (+ layout of :290,:186)
//   CI  cols  S   = IFFT_x(S)                               potentials.py:336 (first half)
//   RI2 rows  t   = exp(i*sigma*Re(IFFT_y(S))*scale)        potentials.py:336-342 + multislice.py:281-282,
//             two slices packed as Re/Im of one complex image (V is real)
//   CP  cols  probe_k * ramp_x * ramp_y -> IFFT_x           multislice.py:221-226
//   RP  rows  IFFT_y * 1/(nx*ny)                            multislice.py:226
//   TW  cols  |fftshift FFT_t(psi - mean_t psi)|^2          tacaw_data.py:94-104
#pragma once
#include "fft_core.cuh"

namespace psb {

enum FftSel { F_NONE = 0, F_FWD = 1, F_INV = 2 };
enum MidSel { M_NONE = 0, M_FULL = 1, M_SEP = 2 };
enum StoreSel { S_PLAIN = 0, S_SHIFT = 1, S_ABS2 = 3, S_TRANSMIT2 = 4 };

struct PassParams {
    const float2* src;          // input images
    float2* dst;                // output images (may alias src)
    long long src_img_stride;   // elements between consecutive source images
    long long dst_img_stride;
    int src_img_mod;            // source image = img % src_img_mod when > 0 (probes shared by frames)
    int nlines;                 // number of lines per image
    int line_len;               // logical line length n
    long long line_stride;      // elements between lines      (rows: n_row, cols: 1)
    long long elem_stride;      // elements along a line       (rows: 1,     cols: row length)
    FftTables tb;
    // M_FULL: psi *= mul[(img / mul_img_div) * mul_img_stride + line*line_stride + p*elem_stride]
    const float2* mul;
    long long mul_img_stride;
    int mul_img_div;
    // M_SEP: psi *= sep_p[img*sep_img_stride_p + p] * sep_l[img*sep_img_stride_l + line]
    const float2* sep_p;
    const float2* sep_l;
    long long sep_img_stride_p, sep_img_stride_l;
    float scale;                // S_PLAIN: output scale; S_TRANSMIT2: V = Re/Im(z)*scale
    float sigma;                // S_TRANSMIT2: t = exp(i*sigma*V)
    float* vout;                // S_TRANSMIT2: optional real potential output (same indexing as dst)
    // S_SHIFT: dst index = (img % probes)*out_stride_probe + (img / probes)*out_stride_frame
    //                      + ((p + n/2) % n)*out_elem_stride + ((line + nlines/2) % nlines)*out_line_stride
    int probes;
    long long out_stride_probe, out_stride_frame, out_elem_stride, out_line_stride;
    // S_SHIFT, slab layout (slab_world > 1; multi-GPU re-sharding): the shifted element rows sq are split
    // into slab_world contiguous blocks (the first `slab_rem` blocks hold slab_base + 1 rows, the rest slab_base), and
    // dst is laid out [block][plane][row in block][line]: block h starts at slab_planes * start_h rows, where a plane
    // is one (layer, frame, probe) image: plane = slab_plane0 + probe*out_stride_probe + frame*out_stride_frame
    // (strides in planes).  Every block is then one contiguous message of the all-to-all, no pack pass.
    int slab_world, slab_base, slab_rem;
    long long slab_planes, slab_plane0;
    // S_ABS2: real output fout[img*dst_img_stride + ((p + n/2) % n)*out_elem_stride + line*out_line_stride]
    float* fout;
    // S_TRANSMIT2: image = frame*pair_count + ml holds V_{2m} + i*V_{2m+1} with m = pair_begin + ml; t of slice s
    //              of that frame goes to dst[(frame*pair_nz + s)*dst_img_stride + ...] (and vout likewise)
    int pair_count, pair_nz, pair_begin;
};

template <int N, int E, int W, bool COLS, bool BLUE, int F1, int MID, int F2, int ST, bool MEANSUB>
struct LinePass {
    static constexpr int T = N / E;
    static constexpr int kThreads = W * T;
#ifdef PSB_LP_MINBLOCKS
    static constexpr int kMinBlocks = PSB_LP_MINBLOCKS;     // tuning experiments only (tools/)
#else
    static constexpr int kMinBlocks = (512 / kThreads) > 0 ? (512 / kThreads) : 1;
#endif
    using Map = SmemMap<N, W, COLS>;
    static constexpr size_t kSmem = Map::kBytes;

    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const PassParams& p) {
        const int tid = cx.tid();
        const int c = COLS ? tid % W : tid / T;
        const int j = COLS ? tid / W : tid % T;
        const int line = cx.bx() * W + c;
        const int img = cx.by();
        const bool line_ok = line < p.nlines;
        const int n = p.line_len;
        float2* sm = reinterpret_cast<float2*>(cx.smem());

        // n == N is a compile-time fact on the direct (power-of-two) path: no per-element bounds tests
        auto in_line = [&](int q) { return BLUE ? q < n : true; };
        const long long estep = (long long)T * p.elem_stride;     // pointer step between registers e, e+1

        const int simg = p.src_img_mod > 0 ? img % p.src_img_mod : img;
        float2 v[E];
        {
            const float2* src = p.src + (long long)simg * p.src_img_stride + (long long)line * p.line_stride
                              + (long long)j * p.elem_stride;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                v[e] = (line_ok && in_line(j + e * T)) ? *src : make_float2(0.f, 0.f);
                src += estep;
            }
        }

        if constexpr (MEANSUB) {
            // mean over the line, tree-reduced through shared memory (T partial sums per line)
            float2 s = make_float2(0.f, 0.f);
#pragma unroll
            for (int e = 0; e < E; ++e) s = cadd(s, v[e]);
            sm[j * W + c] = s;       // plain [j][c] layout, fits: T*W <= N*W
            cx.sync();
            for (int h = T / 2; h > 0; h >>= 1) {
                if (j < h) sm[j * W + c] = cadd(sm[j * W + c], sm[(j + h) * W + c]);
                cx.sync();
            }
            const float inv = 1.0f / (float)n;
            const float2 mean = cscale(sm[c], inv);
            cx.sync();
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (in_line(j + e * T)) v[e] = csub(v[e], mean);
        }

        if constexpr (F1 == F_FWD) dft_line<N, E, W, COLS, -1, BLUE>(cx, v, sm, c, j, p.tb);
        if constexpr (F1 == F_INV) dft_line<N, E, W, COLS, +1, BLUE>(cx, v, sm, c, j, p.tb);

        if constexpr (MID == M_FULL) {
            const float2* mul = p.mul + (long long)(img / p.mul_img_div) * p.mul_img_stride + (long long)line * p.line_stride
                              + (long long)j * p.elem_stride;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (line_ok && in_line(j + e * T)) v[e] = cmul(v[e], __ldg(mul));
                mul += estep;
            }
        }
        if constexpr (MID == M_SEP) {
            const float2 wl = line_ok ? __ldg(&p.sep_l[(long long)img * p.sep_img_stride_l + line]) : make_float2(0.f, 0.f);
            const float2* sp = p.sep_p + (long long)img * p.sep_img_stride_p + j;
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (in_line(j + e * T)) v[e] = cmul(v[e], cmul(__ldg(&sp[e * T]), wl));
        }

        if constexpr (F2 == F_FWD) dft_line<N, E, W, COLS, -1, BLUE>(cx, v, sm, c, j, p.tb);
        if constexpr (F2 == F_INV) dft_line<N, E, W, COLS, +1, BLUE>(cx, v, sm, c, j, p.tb);

        if (!line_ok) return;   // after the last barrier

        if constexpr (ST == S_PLAIN) {
            float2* dst = p.dst + (long long)img * p.dst_img_stride + (long long)line * p.line_stride
                        + (long long)j * p.elem_stride;
            const float sc = p.scale;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (in_line(j + e * T)) *dst = sc == 1.0f ? v[e] : cscale(v[e], sc);
                dst += estep;
            }
        }
        if constexpr (ST == S_SHIFT) {
            const int pr = img % p.probes, fr = img / p.probes;
            const int sl = (line + p.nlines / 2) % p.nlines;
            const int half = n / 2;
            if (p.slab_world > 1) {
                const long long plane = p.slab_plane0 + (long long)pr * p.out_stride_probe + (long long)fr * p.out_stride_frame;
                const int big = p.slab_rem * (p.slab_base + 1);          // rows held by the blocks of slab_base + 1 rows
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int q = j + e * T;
                    if (in_line(q)) {
                        int sq = q + half;
                        if (sq >= n) sq -= n;
                        int start, rows;
                        if (sq < big) {
                            rows = p.slab_base + 1;
                            start = (sq / rows) * rows;
                        } else {
                            rows = p.slab_base;
                            start = big + ((sq - big) / rows) * rows;
                        }
                        p.dst[(p.slab_planes * start + plane * rows + (sq - start)) * p.out_elem_stride + (long long)sl * p.out_line_stride] = v[e];
                    }
                }
            } else {
                float2* dst = p.dst + (long long)pr * p.out_stride_probe + (long long)fr * p.out_stride_frame
                            + (long long)sl * p.out_line_stride;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int q = j + e * T;
                    if (in_line(q)) {
                        int sq = q + half;
                        if (sq >= n) sq -= n;
                        dst[(long long)sq * p.out_elem_stride] = v[e];
                    }
                }
            }
        }
        if constexpr (ST == S_TRANSMIT2) {
            const int fr = img / p.pair_count, m = p.pair_begin + img % p.pair_count;
            const long long base0 = ((long long)fr * p.pair_nz + 2 * m) * p.dst_img_stride + (long long)line * p.line_stride;
            const bool has_b = 2 * m + 1 < p.pair_nz;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int q = j + e * T;
                if (in_line(q)) {
                    const long long o = base0 + (long long)q * p.elem_stride;
                    const float Va = v[e].x * p.scale, Vb = v[e].y * p.scale;
                    float sn, cs;
                    sincosf(p.sigma * Va, &sn, &cs);
                    p.dst[o] = make_float2(cs, sn);
                    if (p.vout) p.vout[o] = Va;
                    if (has_b) {
                        sincosf(p.sigma * Vb, &sn, &cs);
                        p.dst[o + p.dst_img_stride] = make_float2(cs, sn);
                        if (p.vout) p.vout[o + p.dst_img_stride] = Vb;
                    }
                }
            }
        }
        if constexpr (ST == S_ABS2) {
            float* dst = p.fout + (long long)img * p.dst_img_stride + (long long)line * p.out_line_stride;
            const int half = n / 2;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int q = j + e * T;
                if (in_line(q)) {
                    int sq = q + half;
                    if (sq >= n) sq -= n;
                    dst[(long long)sq * p.out_elem_stride] = v[e].x * v[e].x + v[e].y * v[e].y;
                }
            }
        }
    }
};

// Pass identifiers understood by launch_line_pass (host side, line_pass_dispatch.cu)
enum PassKind {
    PASS_R1 = 0,   // rows:  -      * full -> FWD, plain
    PASS_R,        // rows:  INV    * full -> FWD, plain
    PASS_C,        // cols:  FWD    * sep  -> INV, plain
    PASS_CX,       // cols:  FWD, shifted store
    PASS_INV_ROWS, // rows:  INV, plain (scaled)
    PASS_INV_COLS, // cols:  INV, plain (scaled)
    PASS_CP,       // cols:  sep multiply -> INV, plain
    PASS_FWD_ROWS, // rows:  FWD, plain
    PASS_FWD_COLS, // cols:  FWD, plain
    PASS_TW,       // cols:  mean-subtract -> FWD -> |.|^2 shifted
    PASS_RI2,      // rows:  INV, paired transmission epilogue (two slices per complex image)
    PASS_KINDS
};

// Launch `kind` over `n_img` images; chooses the power-of-two size and Bluestein as needed.
// Fills p.tb from the per-device table cache.  Returns 0 or a negative psb error code.
int launch_line_pass(int kind, PassParams p, int n_img, cudaStream_t stream);

}  // namespace psb
