// Thin runtime layer used by the host-side orchestration code: device memory, copies, errors.
// Product build: CUDA runtime.  PSB_EMU build (tests only): plain host memory.
#pragma once
#include "psb_common.cuh"
#include "../../include/pyslice_b200.h"   // psb_status codes

#include <atomic>
#include <mutex>
#include <string>

namespace psb {

void set_error(const std::string& msg);
const char* last_error();
int fail(int code, const std::string& msg);

namespace rt {
void* dev_alloc(size_t bytes);                 // nullptr on failure (error text set)
void dev_free(void* p);
int h2d(void* dst, const void* src, size_t bytes, cudaStream_t s);   // synchronous w.r.t. host buffer
int zero(void* dst, size_t bytes, cudaStream_t s);
int device();                                  // current device ordinal
int check(const char* what);                   // cudaGetLastError -> psb code
int sm_count();

// One-time set-up that CUDA keeps per device (cudaFuncSetAttribute, ...): `run(f)` calls f() the first time the object is
// used on the CURRENT device and remembers success per device ordinal, under a lock, so a process that drives several
// GPUs (or several host threads) never skips the set-up on a device that has not seen it.
struct PerDeviceOnce {
    std::atomic<unsigned long long> done{0};
    std::mutex mu;
    template <class F>
    int run(F&& f) {
        const int d = device() & 63;
        if ((done.load(std::memory_order_acquire) >> d) & 1ull) return PSB_OK;
        std::lock_guard<std::mutex> lk(mu);
        if ((done.load(std::memory_order_relaxed) >> d) & 1ull) return PSB_OK;
        const int rc = f();
        if (rc == PSB_OK) done.fetch_or(1ull << d, std::memory_order_release);
        return rc;
    }
};
}  // namespace rt

}  // namespace psb
