/* pyslice_b200 -- C ABI of the B200 (sm_100a) multislice + TACAW engine (libpsb.so).
 *
 * The reference (h-walk/PySlice) has no FFI: its hot path is a chain of torch calls inside Python
 * classes.  These entry points are what a binding for that path replaces; each one cites the
 * reference lines whose arithmetic it performs (paths relative to the reference checkout).  See
 * INTEGRATION.md for the ctypes stubs a PySlice maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative psb_status; psb_last_error() gives the
 *     text for the calling thread;
 *   - all array arguments are caller-owned DEVICE pointers on the current CUDA device
 *     (torch: tensor.data_ptr()); `stream` is a cudaStream_t passed as void* (NULL = default);
 *   - complex64 arrays are interleaved (re, im) float pairs; images are row-major (nx, ny), ny fastest;
 *   - calls are asynchronous on `stream` unless stated; no callbacks, no exceptions, no host
 *     fallback -- without a CUDA device every call fails with PSB_ERR_CUDA.
 */
#ifndef PYSLICE_B200_H
#define PYSLICE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PSB_OK = 0,
    PSB_ERR_INVALID = -1,
    PSB_ERR_UNSUPPORTED = -2,
    PSB_ERR_CUDA = -3,
    PSB_ERR_NOMEM = -4
} psb_status;

typedef struct { float re, im; } psb_c64;

int psb_version(void);
const char* psb_last_error(void);
/* number of SMs of the current device (148 on B200); used by the host to size frame batches */
int psb_sm_count(void);
/* drop cached FFT tables of all devices */
void psb_release_tables(void);
/* kernels launched by this library in this process so far (benchmark bookkeeping) */
long long psb_launch_count(void);
/* diagnostic switch between kernel generations (all CUDA; used by microbenchmarks and A/B parity tests):
 * 0 = generic line-pass kernels, anything else = fused persistent kernels (default). */
void psb_set_fast_path(int level);
/* CUDA-graph replay of repeated launch sequences (default on; environment PSB_GRAPHS=0 or psb_set_graph_mode(0) turn it off):
 * psb_propagate_ex / psb_propagate / psb_propagate_phase and psb_build_transmission / psb_build_phase called again with the
 * same buffers and sizes replay the ~1000 launches of a batch as one graph launch.  Results are identical either way. */
void psb_set_graph_mode(int on);
/* structure-factor algorithm of psb_build_transmission / psb_build_phase: 0 = automatic (default: the direct sum, or the
 * 1-D NUFFT of csrc/sf_nufft.cu for dense slices on 1024-point grids), 1 = direct sum always, 2 = NUFFT wherever it is
 * implemented (nx = 512 or 1024).  Both build the same spectrum to float32 round-off. */
void psb_set_sf_mode(int mode);

/* ---- atom -> slice binning: src/multislice/potentials.py:297-317 (+ bounds :304-305) -------------
 * positions (F, A, 3) float64; type_idx (A) dense type index in [0, ntypes);
 * lo, hi (nz) float64 slice bounds evaluated by the host with the reference's expressions.
 * Outputs: offsets (F, nz*ntypes+1) exclusive segment offsets (segment = slice*ntypes + type);
 *          atom_list / ux / uy (F, 2*A): atom index and frac(x/Lx), frac(y/Ly) as 32-bit fixed point,
 *          grouped by segment, ascending atom index inside a segment (deterministic).
 * seg_scratch: int32 workspace of F*(4*A + nz*ntypes) elements (segment ids, unsorted lists, fill cursors).
 * lx_eff = nx*dx, ly_eff = ny*dy. */
int psb_bin_atoms(const double* positions, const int32_t* type_idx, int n_frames, int n_atoms, int ntypes,
                  int nz, const double* lo, const double* hi, double dz, double lx_eff, double ly_eff,
                  int32_t* seg_scratch, int32_t* offsets, int32_t* atom_list, uint32_t* ux, uint32_t* uy,
                  void* stream);

/* ---- projected potential -> transmission: potentials.py:319-342 and multislice.py:281-282 --------
 * formfactors (ntypes, nx, ny) float32 = Kirkland f_e on the fftfreq grid (host table, potentials.py:86-96).
 * t_out (F, nz, nx, ny) complex64 = exp(i*sigma*V), V = Re IFFT2(S) * scale  with scale = 1/(dx^2 dy^2)
 * (the 1/(nx*ny) of the inverse FFT is applied internally).  v_out (same shape, float32) optional.
 * scratch: complex64 workspace of scratch_elems >= nx*ny elements; the stack is built in chunks of
 * slice-pair images that fit it (size it to stay L2-resident, e.g. 32 MiB).  V is real, so two slices share one complex inverse FFT and only
 * half of each spectrum is summed (the Hermitian part -- exactly what the reference's Re() keeps). */
int psb_build_transmission(const int32_t* offsets, const uint32_t* ux, const uint32_t* uy, int n_frames,
                           int n_atoms, int nz, int ntypes, int nx, int ny, const float* formfactors,
                           float scale, float sigma, psb_c64* t_out, float* v_out, psb_c64* scratch,
                           long long scratch_elems, void* stream);

/* The same build with the stack kept as float32 phases sigma*V[f, z, x, y] (4 B per pixel and slice instead of 8):
 * multislice.py:281-282 evaluates exp(i*sigma*V) per slice anyway, psb_propagate_phase does so inside its fused row
 * pass.  Worth it when few probes share a frame's stack (plane-wave runs: written once, read once).  Grids with fused
 * kernels only (psb_phase_format_supported), PSB_ERR_UNSUPPORTED otherwise. */
int psb_build_phase(const int32_t* offsets, const uint32_t* ux, const uint32_t* uy, int n_frames,
                    int n_atoms, int nz, int ntypes, int nx, int ny, const float* formfactors,
                    float scale, float sigma, float* phase_out, psb_c64* scratch, long long scratch_elems, void* stream);
int psb_phase_format_supported(int nx, int ny);

/* t = exp(i*sigma*V) for a user-supplied real potential (Propagate() on a Potential object) */
int psb_transmission_from_potential(const float* v, psb_c64* t, long long n, float sigma, void* stream);

/* ---- generic batched 2-D FFT, unnormalised forward / inverse*scale: torch.fft.fft2/ifft2 call sites
 * multislice.py:124,188,218  (probe construction, defocus).  In place allowed. */
int psb_fft2(const psb_c64* src, psb_c64* dst, int batch, int nx, int ny, int inverse, float scale, void* stream);

/* ---- shifted probes: multislice.py:216-231.  out[p] = ifft2(base_k * ramp_x[p][:,None] * ramp_y[p][None,:]) */
int psb_shift_probes(const psb_c64* base_k, const psb_c64* ramp_x, const psb_c64* ramp_y, int n_probes,
                     int nx, int ny, psb_c64* out, void* stream);

/* ---- multislice propagation + exit FFT: multislice.py:278-294, calculators.py:285-290,185-186 -----
 * probes (P, nx, ny); t (F, nz, nx, ny); prop_x (nx), prop_y (ny): separable Fresnel propagator
 * exp(-i*pi*lambda*dz*k^2) with the 1/(nx*ny) of the FFT pair folded in by the host.
 * psi_work (F*P, nx, ny) scratch, image index = frame*P + probe.
 * mode 0: exit waves in real space are left in psi_work (Propagate()).
 * mode 1: wf_out[layer*stride_layer + probe*stride_probe + frame*stride_frame + kx'*ny + ky'] receives
 *         fftshift(fft2(psi)) after the transmission of every `layer_every`-th slice (0: exit only;
 *         the exit wave is always the last layer). */
int psb_propagate(const psb_c64* probes, const psb_c64* t, int n_frames, int n_probes, int nz, int nx, int ny,
                  const psb_c64* prop_x, const psb_c64* prop_y, psb_c64* psi_work, int mode, psb_c64* wf_out,
                  long long stride_probe, long long stride_frame, long long stride_layer, int layer_every,
                  void* stream);

/* psb_propagate on a phase stack from psb_build_phase; t0 (n_frames, nx, ny) is scratch for exp(i*phase) of slice 0 */
int psb_propagate_phase(const psb_c64* probes, const float* phase, psb_c64* t0, int n_frames, int n_probes, int nz, int nx,
                        int ny, const psb_c64* prop_x, const psb_c64* prop_y, psb_c64* psi_work, int mode, psb_c64* wf_out,
                        long long stride_probe, long long stride_frame, long long stride_layer, int layer_every,
                        void* stream);

/* The general form of the two calls above, one descriptor instead of 18 positional arguments, plus two output modes the
 * multi-GPU run and the STEM detectors need.  `struct_bytes` = sizeof(psb_propagate_desc) (checked).
 *   t / phase      exactly one is non-NULL: complex64 stack, or float32 phase stack (then t0_scratch (F, nx, ny) is required)
 *   mode 0, 1      as psb_propagate
 *   slab_world > 1 (mode 1): wf_out is laid out for the frames -> kx-rows all-to-all that precedes the time FFT
 *                  (replaces the pack pass around calculators.py:285-290,185-186): the shifted kx rows are
 *                  split into slab_world contiguous blocks (the first nx % slab_world blocks hold one row more), and
 *                  wf_out[block h][layer][frame][probe][row in block][ky'] with slab_layers x slab_frames x slab_probes
 *                  planes per block; this call fills frames [frame0, frame0 + n_frames) and probes [probe0, probe0 + n_probes).
 *                  Each block is one contiguous message.  stride_* are ignored.
 *   mode 2         detector sums only (haadf_data.py:43-65 without the cube): at every layer tap the shifted k-space image
 *                  goes to det_scratch (n_frames*n_probes, nx, ny) and det_out[layer*det_stride_layer +
 *                  (probe0 + probe)*det_stride_probe + frame0 + frame] = sum_k |psi_k| * det_mask[kx', ky'] (float64). */
typedef struct psb_propagate_desc {
    int struct_bytes;
    int mode, layer_every;
    int n_frames, n_probes, nz, nx, ny;
    const psb_c64* probes;
    const psb_c64* t;
    const float* phase;
    psb_c64* t0_scratch;
    const psb_c64* prop_x;
    const psb_c64* prop_y;
    psb_c64* psi_work;
    psb_c64* wf_out;
    long long stride_probe, stride_frame, stride_layer;
    int slab_world, slab_layers, slab_frames, slab_probes, frame0, probe0;
    const float* det_mask;
    double* det_out;
    long long det_stride_layer, det_stride_probe;
    psb_c64* det_scratch;
    void* stream;
} psb_propagate_desc;
int psb_propagate_ex(const psb_propagate_desc* desc);

/* ---- TACAW: tacaw_data.py:89-104.  intensity[p, w, pix] = |fftshift_t FFT_t(psi - mean_t psi)|^2
 * wf element (p, f, pix) at wf[p*stride_probe + f*stride_frame + pix]; intensity (P, T, npix) float32. */
int psb_tacaw_intensity(const psb_c64* wf, long long stride_probe, long long stride_frame, int n_probes,
                        int n_frames, long long npix, float* intensity, void* stream);

/* ---- reducers: tacaw_data.py:124,137,198,211,277,293; haadf_data.py:60 -----------------------------
 * out[row] = sum_pix in[row*row_stride + pix] * mask[pix] (mask may be NULL), float64 results */
int psb_sum_pixels(const float* in, const float* mask, int rows, long long row_stride, long long npix,
                   double* out, void* stream);
/* same with |z| of a complex input (HAADF detector sum) */
int psb_sum_abs_pixels(const psb_c64* in, const float* mask, int rows, long long row_stride, long long npix,
                       double* out, void* stream);
/* out[g, pix] = sum_t in[g, t, pix] */
int psb_sum_frames(const float* in, int groups, int n_frames, long long npix, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYSLICE_B200_H */
