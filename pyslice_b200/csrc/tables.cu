#include "tables.h"

#include <cmath>
#include <complex>
#include <map>
#include <mutex>
#include <vector>

namespace psb {

namespace {
struct Entry {
    float2* tw = nullptr;
    float2* chirp = nullptr;
    float2* bhat = nullptr;
    int N = 0;
    bool blue = false;
};
std::mutex g_mu;
std::map<std::pair<int, int>, Entry> g_cache;    // (device, n) -> tables
std::map<std::pair<int, int>, float2*> g_tw;     // (device, N) -> twiddles

typedef std::complex<double> cd;
const double kPi = 3.14159265358979323846264338327950288;

// exact-argument unit phasor exp(sign*2*pi*i*num/den) with num reduced mod den in integers
cd unit(long long num, long long den, int sign) {
    num %= den;
    if (num < 0) num += den;
    double a = 2.0 * kPi * (double)num / (double)den;
    return cd(std::cos(a), sign * std::sin(a));
}

void host_fft(std::vector<cd>& a) {   // in-place radix-2, forward, power-of-two length
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                cd w = unit((long long)k, (long long)len, -1);
                cd u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

// staged twiddle table of the plan used for line size N (layout: fft_core.cuh, twiddle_offset)
template <int N>
std::vector<cd> staged_for() {
    constexpr int E = elems_for(N);
    using P = Plan<N, E>;
    std::vector<cd> out;
    int ns = 1;
    for (int s = 0; s < P::S; ++s) {
        const int R = P::R[s];
        if (ns > 1)
            for (int t = 1; t < R; ++t)
                for (int k = 0; k < ns; ++k) out.push_back(unit((long long)k * t, (long long)ns * R, -1));
        ns *= R;
    }
    if (out.empty()) out.push_back(cd(1, 0));
    return out;
}
std::vector<cd> staged_twiddles(int N) {
    switch (N) {
        case 16: return staged_for<16>();
        case 32: return staged_for<32>();
        case 64: return staged_for<64>();
        case 128: return staged_for<128>();
        case 256: return staged_for<256>();
        case 512: return staged_for<512>();
        case 1024: return staged_for<1024>();
        case 2048: return staged_for<2048>();
        case 4096: return staged_for<4096>();
        case 8192: return staged_for<8192>();
    }
    return std::vector<cd>(1, cd(1, 0));
}

float2* upload(const std::vector<cd>& h, cudaStream_t s) {
    std::vector<float2> f(h.size());
    for (size_t i = 0; i < h.size(); ++i) f[i] = make_float2((float)h[i].real(), (float)h[i].imag());
    float2* d = (float2*)rt::dev_alloc(f.size() * sizeof(float2));
    if (!d) return nullptr;
    if (rt::h2d(d, f.data(), f.size() * sizeof(float2), s) != PSB_OK) return nullptr;
    return d;
}
}  // namespace

int fft_size_for(int n, bool* bluestein) {
    if (n < 1) return 0;
    if ((n & (n - 1)) == 0) {
        *bluestein = false;
        int N = n < 16 ? 16 : n;       // tiny powers of two ride in a 16-point line via Bluestein
        if (n < 16) *bluestein = true;
        return N <= kMaxLine ? N : 0;
    }
    *bluestein = true;
    int N = 16;
    while (N < 2 * n - 1) N <<= 1;
    return N <= kMaxLine ? N : 0;
}

int get_fft_tables(int n, FftTables* out, int* N_out, bool* blue_out, cudaStream_t s) {
    bool blue = false;
    int N = fft_size_for(n, &blue);
    if (N == 0) return fail(PSB_ERR_UNSUPPORTED, "FFT length " + std::to_string(n) + " not supported (max 8192, or 4096 if not a power of two)");
    std::lock_guard<std::mutex> lk(g_mu);
    const int dev = rt::device();
    auto key = std::make_pair(dev, n);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
        Entry e;
        e.N = N;
        e.blue = blue;
        auto tk = std::make_pair(dev, N);
        auto tt = g_tw.find(tk);
        if (tt == g_tw.end()) {
            std::vector<cd> tw = staged_twiddles(N);
            float2* d = upload(tw, s);
            if (!d) return PSB_ERR_NOMEM;
            tt = g_tw.emplace(tk, d).first;
        }
        e.tw = tt->second;
        if (blue) {
            std::vector<cd> chirp(n), b(N, cd(0, 0));
            for (long long p = 0; p < n; ++p) chirp[p] = unit(p * p, 2LL * n, -1);   // exp(-i*pi*p^2/n)
            for (long long m = 0; m < n; ++m) {
                cd v = std::conj(chirp[m]);
                b[m] = v;
                if (m) b[N - m] = v;
            }
            host_fft(b);
            for (auto& x : b) x /= (double)N;
            e.chirp = upload(chirp, s);
            e.bhat = upload(b, s);
            if (!e.chirp || !e.bhat) return PSB_ERR_NOMEM;
        }
        it = g_cache.emplace(key, e).first;
    }
    out->tw = it->second.tw;
    out->chirp = it->second.chirp;
    out->bhat = it->second.bhat;
    out->n = n;
    *N_out = it->second.N;
    *blue_out = it->second.blue;
    return PSB_OK;
}

void free_all_tables() {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_cache) {
        rt::dev_free(kv.second.chirp);
        rt::dev_free(kv.second.bhat);
    }
    for (auto& kv : g_tw) rt::dev_free(kv.second);
    g_cache.clear();
    g_tw.clear();
}

}  // namespace psb
