#include "psb_rt.h"

#include <mutex>

namespace psb {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }
int fail(int code, const std::string& msg) {
    set_error(msg);
    return code;
}

namespace rt {
#ifndef PSB_EMU
void* dev_alloc(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
void dev_free(void* p) { if (p) cudaFree(p); }
int h2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("h2d: ") + cudaGetErrorString(e));
    return PSB_OK;
}
int zero(void* dst, size_t bytes, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(dst, 0, bytes, s);
    if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("memset: ") + cudaGetErrorString(e));
    return PSB_OK;
}
int device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}
int check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PSB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return PSB_OK;
}
int sm_count() {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device());
    return n > 0 ? n : 148;
}
#else
void* dev_alloc(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
void dev_free(void* p) { std::free(p); }
int h2d(void* dst, const void* src, size_t bytes, cudaStream_t) { std::memcpy(dst, src, bytes); return PSB_OK; }
int zero(void* dst, size_t bytes, cudaStream_t) { std::memset(dst, 0, bytes); return PSB_OK; }
int device() { return 0; }
int check(const char*) { return PSB_OK; }
int sm_count() { return 4; }
#endif
}  // namespace rt
}  // namespace psb
