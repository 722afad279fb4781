// Thin runtime layer used by the host-side orchestration code: device memory, copies, errors.
// Product build: CUDA runtime.  PSB_EMU build (tests only): plain host memory.
#pragma once
#include "psb_common.cuh"
#include "../../include/pyslice_b200.h"   // psb_status codes

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
}  // namespace rt

}  // namespace psb
