// Per-device cache of FFT tables (twiddles, Bluestein chirps), built on the host in float64 and
// rounded once to float32 -- no recurrences, so every table entry is correctly rounded.
#pragma once
#include "fft_core.cuh"
#include "psb_rt.h"

namespace psb {

// smallest supported power of two N for a logical length n: n itself if it is a power of two,
// else the first power of two >= 2n-1 (Bluestein).  Returns 0 if unsupported.
int fft_size_for(int n, bool* bluestein);

// Device tables for logical length n (cached per device).  Returns PSB_OK or an error code.
int get_fft_tables(int n, FftTables* out, int* N_out, bool* blue_out, cudaStream_t s);

void free_all_tables();

}  // namespace psb
