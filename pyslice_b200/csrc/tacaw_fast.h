// Host entry points of the tiled mixed-radix TACAW time transform (tacaw_fast.cu).
#pragma once
#include "psb_common.cuh"

namespace psb {

// true when n_frames = 2^a 3^b 5^c and a tile of at least 4 pixels x n_frames fits shared memory
bool tacaw_fast_supported(int n_frames);
// intensity[p, w, k] = |fftshift_t FFT_t(wf[p, t, k] - mean_t)|^2; wf strides in elements, pixels contiguous
int launch_tacaw_fast(const float2* wf, long long stride_probe, long long stride_frame, int n_probes, int n_frames,
                      long long npix, float* intensity, cudaStream_t s);
void tacaw_fast_release();

}  // namespace psb
