// Host entry points of the fused slice-step kernels (fast_path.cu).
#pragma once
#include "psb_common.cuh"

namespace psb {

// true when both passes of a slice step have a fused kernel for this grid (and the fast path is enabled)
bool fast_slice_supported(int nx, int ny);
void fast_path_enable(int on);

// psi[x, ky] -> FFT_y( t[x, y] * IFFT_y( py[ky] * psi[x, ky] ) ) in place over n_img contiguous (nx, ny) images;
// image i uses the transmission slice t_slice + (i / probes) * t_frame_stride.
int launch_fast_rows(float2* psi, int n_img, int nx, int ny, const float2* t_slice, long long t_frame_stride,
                     int probes, const float2* py, cudaStream_t s);
// psi[x, ky] -> IFFT_x( px[kx] * FFT_x( psi[x, ky] ) ) in place (unnormalised; px carries the 1/(nx*ny)).
int launch_fast_cols(float2* psi, int n_img, int nx, int ny, const float2* px, cudaStream_t s);

}  // namespace psb
