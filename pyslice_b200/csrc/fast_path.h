// Host entry points of the fused slice-step kernels (fast_path.cu).
#pragma once
#include "psb_common.cuh"

namespace psb {

// true when both passes of a slice step have a fused kernel for this grid (and the fast path is enabled)
bool fast_slice_supported(int nx, int ny);
// level 0: generic line passes only; 1 (default): fused slice-step / transmission kernels and pipelined structure
// factor (sf_fast.cu)
void fast_path_enable(int level);
bool fast_path_enabled();
int fast_path_level();

// psi[x, ky] -> FFT_y( t[x, y] * IFFT_y( psi[x, ky] ) ) in place over n_img contiguous (nx, ny) images;
// image i uses the transmission slice t_slice + (i / probes) * t_frame_stride.
int launch_fast_rows(float2* psi, int n_img, int nx, int ny, const float2* t_slice, long long t_frame_stride,
                     int probes, cudaStream_t s);
// psi[x, ky] -> IFFT_x( px[kx] * py[ky] * FFT_x( psi[x, ky] ) ) in place (unnormalised; px carries the 1/(nx*ny)).
int launch_fast_cols(float2* psi, int n_img, int nx, int ny, const float2* px, const float2* py, cudaStream_t s);

// potential build (potentials.py:336-342): in-place inverse column transform of n_img slice-pair spectra, then
// inverse row transform with the transmission epilogue t = exp(i*sigma*scale*Re/Im(.)) for the two slices of a pair
int launch_fast_cols_inverse(float2* imgs, int n_img, int nx, int ny, cudaStream_t s);
int launch_fast_rows_transmit(float2* pairs, int n_img, int nx, int ny, float scale, float sigma, float2* t_out,
                              float* v_out, int pair_count, int pair_nz, int pair_begin, cudaStream_t s);

// the same two passes with the transmission stack kept as float32 phases sigma*V (t = exp(i*phase) is evaluated by the
// slice step as it multiplies): phase_out / phase_slice are indexed like t_out / t_slice
int launch_fast_rows_phase(float2* psi, int n_img, int nx, int ny, const float* phase_slice, long long t_frame_stride,
                           int probes, cudaStream_t s);
int launch_fast_rows_phase_out(float2* pairs, int n_img, int nx, int ny, float scale, float sigma, float* phase_out,
                               int pair_count, int pair_nz, int pair_begin, cudaStream_t s);

// structure-factor sum of slice pairs [pair_begin, pair_begin + pair_count) of nf frames into out (nf, pair_count, nx, ny)
// (sf_fast.cu: precomputed phase tables + TMA-fed packed-FMA tiles); any grid size
// sf_fast_prepare gathers the form-factor table once per psb_build_transmission call; launch_sf_fast runs per chunk
// `s` receives the launches; the workspaces are kept per (device, `owner` stream) -- the caller's stream, which differs from
// `s` only while the sequence is being recorded into a graph on the library's capture stream
int sf_fast_prepare(const float* ff, int ntypes, int nx, int ny, cudaStream_t s, cudaStream_t owner);
int launch_sf_fast(const int* offsets, const unsigned int* ux, const unsigned int* uy, int cap, int nz, int ntypes, int nx,
                   int ny, int pair_begin, int pair_count, int nf, float2* out, cudaStream_t s, cudaStream_t owner);
void sf_fast_release();
bool sf_fast_supported(int ntypes);

// structure factor of dense slices through a 1-D NUFFT along x (sf_nufft.cu): same output as launch_sf_fast.
// sf_nufft_wanted: the grid is supported and the slices are dense enough for it to beat the direct sum (or the mode forces it)
void sf_mode_set(int mode);      // 0 auto (default), 1 direct sum always, 2 NUFFT wherever supported
int sf_mode();
bool sf_nufft_supported(int ntypes, int nx, int ny);
bool sf_nufft_wanted(int ntypes, int nx, int ny, int n_atoms, int nz);
// sf_nufft_prepare: once per build call over all n_frames (offsets / ux / uy of frame 0); launch_sf_nufft: per chunk
int sf_nufft_prepare(const int* offsets, const unsigned int* ux, const unsigned int* uy, int cap, int nz, int ntypes, int nx,
                     int ny, int n_frames, const float* ff, cudaStream_t s, cudaStream_t owner);
int launch_sf_nufft(const int* offsets, const unsigned int* ux, const unsigned int* uy, int cap, int nz, int ntypes, int nx,
                    int ny, int n_frames, int frame0, int nf, int pair_begin, int pair_count, const float* ff, float2* out,
                    cudaStream_t s, cudaStream_t owner);
void sf_nufft_release();      // the pipelined kernel stages at most 64 atom types; beyond that the generic kernel runs

}  // namespace psb
