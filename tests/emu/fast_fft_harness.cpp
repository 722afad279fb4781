// TEST INFRASTRUCTURE (never loaded by the package): runs the fast-path line transform of
// pyslice_b200/csrc/fast_fft.cuh on the host, one std::thread per CUDA thread of a line, so the CPU test
// suite can check its stage index arithmetic and twiddle selection against numpy.
// build: g++ -std=c++20 -O2 -fPIC -pthread -DPSB_EMU -shared -o libfastfft_emu.so fast_fft_harness.cpp
#include "../../pyslice_b200/csrc/fast_fft.cuh"

#include <barrier>
#include <cmath>
#include <thread>
#include <vector>

using namespace psb;

namespace {
struct HostXchg {
    fast::cpx* a;
    fast::cpx* b;
    std::barrier<>* bar;
    fast::cpx* buf(int i) const { return (i & 1) ? b : a; }
    int at(int q) const { return q; }
    void after_store(int) const { bar->arrive_and_wait(); }
    void after_load(int) const {}          // alternating buffers: the next exchange's barrier orders the reuse
    void mid_sync(int) const { bar->arrive_and_wait(); }
};

template <int N>
std::vector<float2> staged_table() {
    const double pi = 3.14159265358979323846;
    std::vector<float2> t;
    auto w = [&](long long num, long long den) {
        double a = -2.0 * pi * (double)(num % den) / (double)den;
        return make_float2((float)std::cos(a), (float)std::sin(a));
    };
    if (N == 256) {
        for (int tt = 1; tt < 16; ++tt)
            for (int k = 0; k < 16; ++k) t.push_back(w((long long)k * tt, 256));
    } else if (N == 512) {
        for (int k = 0; k < 16; ++k) t.push_back(w(k, 32));
        for (int tt = 1; tt < 16; ++tt)
            for (int k = 0; k < 32; ++k) t.push_back(w((long long)k * tt, 512));
    } else if (N == 1024) {
        for (int tt = 1; tt < 4; ++tt)
            for (int k = 0; k < 16; ++k) t.push_back(w((long long)k * tt, 64));
        for (int tt = 1; tt < 16; ++tt)
            for (int k = 0; k < 64; ++k) t.push_back(w((long long)k * tt, 1024));
    } else {
        for (int tt = 1; tt < 8; ++tt)
            for (int k = 0; k < 16; ++k) t.push_back(w((long long)k * tt, 128));
        for (int tt = 1; tt < 16; ++tt)
            for (int k = 0; k < 128; ++k) t.push_back(w((long long)k * tt, 2048));
    }
    return t;
}

template <int N, int DIR>
void run_line(const float2* in, float2* out, int reps) {
    constexpr int T = N / 16;
    std::vector<float2> xa(N), xb(N);
    std::vector<float2> table = staged_table<N>();
    std::barrier<> bar(T);
    std::vector<std::thread> th;
    for (int j = 0; j < T; ++j)
        th.emplace_back([&, j] {
            fast::Twiddles<N> tw;
            tw.load(table.data(), j);
            fast::cpx v[16];
            for (int e = 0; e < 16; ++e) v[e] = in[j + e * T];
            HostXchg x{xa.data(), xb.data(), &bar};
            for (int r = 0; r < reps; ++r) {
                fast::cpx w[16];
                fast::line_fft<N, DIR>([&](int e) { return v[e]; }, [&](int e, fast::cpx a) { w[e] = a; }, tw, j, x,
                                       r * fast::exchanges<N>());
                for (int e = 0; e < 16; ++e) v[e] = w[e];
            }
            for (int e = 0; e < 16; ++e) out[j + e * T] = v[e];
        });
    for (auto& t : th) t.join();
}
}  // namespace

extern "C" int fast_fft_line(int N, int dir, const float* in, float* out, int reps) {
    const float2* i2 = reinterpret_cast<const float2*>(in);
    float2* o2 = reinterpret_cast<float2*>(out);
    if (N == 256 && dir < 0) run_line<256, -1>(i2, o2, reps);
    else if (N == 256) run_line<256, +1>(i2, o2, reps);
    else if (N == 512 && dir < 0) run_line<512, -1>(i2, o2, reps);
    else if (N == 512) run_line<512, +1>(i2, o2, reps);
    else if (N == 1024 && dir < 0) run_line<1024, -1>(i2, o2, reps);
    else if (N == 1024) run_line<1024, +1>(i2, o2, reps);
    else if (N == 2048 && dir < 0) run_line<2048, -1>(i2, o2, reps);
    else if (N == 2048) run_line<2048, +1>(i2, o2, reps);
    else return -1;
    return 0;
}
