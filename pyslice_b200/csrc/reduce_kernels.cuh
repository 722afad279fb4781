// Small reductions behind the TACAW / HAADF reducers (reference src/postprocessing/tacaw_data.py:109-299,
// haadf_data.py:60): sums of the float32 intensity cube over pixels or over frequency.
#pragma once
#include "psb_common.cuh"

namespace psb {

struct SumPixParams {
    const float* in;        // real input rows (or nullptr)
    const float2* cin;      // complex input rows: |z| is summed (or nullptr)
    const float* mask;      // optional per-pixel weight
    long long row_stride;   // elements between rows
    long long npix;
    double* out;            // one value per row: out[(row % row_mod) * out_stride_mod + (row / row_mod) * out_stride_div]
    int row_mod;            // 0: out[row]
    long long out_stride_mod, out_stride_div;
};

// out[row] = sum_pix in[row, pix] * mask[pix]   -- one CTA per row, float64 accumulation
struct SumPixels {
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 1;
    static constexpr size_t kSmem = kThreads * sizeof(double);
    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const SumPixParams& p) {
        const long long row = cx.bx();
        const int t = cx.tid();
        double acc = 0.0;
        for (long long i = t; i < p.npix; i += kThreads) {
            float v;
            if (p.cin) {
                const float2 z = p.cin[row * p.row_stride + i];
                v = sqrtf(z.x * z.x + z.y * z.y);
            } else {
                v = p.in[row * p.row_stride + i];
            }
            acc += p.mask ? (double)(v * p.mask[i]) : (double)v;
        }
        double* sm = reinterpret_cast<double*>(cx.smem());
        sm[t] = acc;
        cx.sync();
        for (int h = kThreads / 2; h > 0; h >>= 1) {
            if (t < h) sm[t] += sm[t + h];
            cx.sync();
        }
        if (t == 0) {
            if (p.row_mod > 0) p.out[(row % p.row_mod) * p.out_stride_mod + (row / p.row_mod) * p.out_stride_div] = sm[0];
            else p.out[row] = sm[0];
        }
    }
};

struct SumRowsParams {
    const float* in;    // (G, T, npix)
    int T;
    long long npix;
    float* out;         // (G, npix)
};

// out[g, pix] = sum_t in[g, t, pix]
struct SumRows {
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 1;
    template <class Ctx>
    static PSB_D void run(const Ctx& cx, const SumRowsParams& p) {
        const long long pix = (long long)cx.bx() * kThreads + cx.tid();
        if (pix >= p.npix) return;
        const float* src = p.in + (long long)cx.by() * p.T * p.npix + pix;
        double acc = 0.0;
        for (int t = 0; t < p.T; ++t) acc += (double)src[(long long)t * p.npix];
        p.out[(long long)cx.by() * p.npix + pix] = (float)acc;
    }
};

}  // namespace psb
