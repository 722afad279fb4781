// C ABI of libpsb.so: argument checking and kernel orchestration (see include/pyslice_b200.h).
#include "../../include/pyslice_b200.h"

#include "fast_path.h"
#include "graph_cache.h"
#include "tacaw_fast.h"
#include "line_pass.cuh"
#include "potential_kernels.cuh"
#include "psb_rt.h"
#include "reduce_kernels.cuh"
#include "tables.h"

using namespace psb;

namespace {
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline float2* f2(psb_c64* p) { return reinterpret_cast<float2*>(p); }
inline const float2* f2(const psb_c64* p) { return reinterpret_cast<const float2*>(p); }

template <class K, class P>
int go(dim3 grid, size_t smem, cudaStream_t s, const P& p, const char* what) {
    cudaError_t e = launch<K>(grid, smem, s, p);
    if (e != cudaSuccess) {
#ifndef PSB_EMU
        return fail(PSB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
#else
        return PSB_ERR_CUDA;
#endif
    }
    return PSB_OK;
}

PassParams base_params() {
    PassParams p;
    std::memset(&p, 0, sizeof(p));
    p.scale = 1.0f;
    p.mul_img_div = 1;
    p.probes = 1;
    return p;
}

// rows-oriented pass over (n_img, nx, ny) images: lines are the nx rows of length ny
void rows_geometry(PassParams& p, int nx, int ny) {
    p.nlines = nx;
    p.line_len = ny;
    p.line_stride = ny;
    p.elem_stride = 1;
}
// cols-oriented pass: lines are the ny columns of length nx
void cols_geometry(PassParams& p, int nx, int ny) {
    p.nlines = ny;
    p.line_len = nx;
    p.line_stride = 1;
    p.elem_stride = ny;
}
}  // namespace

extern "C" {

int psb_version(void) { return 100; }
const char* psb_last_error(void) { return last_error(); }
int psb_sm_count(void) { return rt::sm_count(); }
void psb_release_tables(void) {
#ifndef PSB_EMU
    graph_cache_release();        // cached graphs hold pointers into the tables freed below
#endif
    free_all_tables();
#ifndef PSB_EMU
    sf_fast_release();
    tacaw_fast_release();
#endif
}
long long psb_launch_count(void) { return launch_counter(); }
void psb_set_graph_mode(int on) {
#ifndef PSB_EMU
    graph_mode_set(on);
#else
    (void)on;
#endif
}
void psb_set_sf_mode(int mode) {
#ifndef PSB_EMU
    sf_mode_set(mode);
#else
    (void)mode;
#endif
}
void psb_set_fast_path(int level) {
#ifndef PSB_EMU
    fast_path_enable(level);
#else
    (void)level;
#endif
}

int psb_bin_atoms(const double* positions, const int32_t* type_idx, int n_frames, int n_atoms, int ntypes,
                  int nz, const double* lo, const double* hi, double dz, double lx_eff, double ly_eff,
                  int32_t* seg_scratch, int32_t* offsets, int32_t* atom_list, uint32_t* ux, uint32_t* uy,
                  void* stream) {
    if (!positions || !type_idx || !lo || !hi || !seg_scratch || !offsets || !atom_list || !ux || !uy)
        return fail(PSB_ERR_INVALID, "psb_bin_atoms: null pointer");
    if (n_frames < 0 || n_atoms < 0 || ntypes < 1 || nz < 1 || !(dz > 0) || !(lx_eff > 0) || !(ly_eff > 0))
        return fail(PSB_ERR_INVALID, "psb_bin_atoms: bad sizes");
    if (n_frames > 65535 || (long long)nz * ntypes > 0x7fffffffLL)
        return fail(PSB_ERR_UNSUPPORTED, "psb_bin_atoms: too many frames per call (max 65535)");
    cudaStream_t s = as_stream(stream);
    const int nseg = nz * ntypes;
    int rc = rt::zero(offsets, (size_t)n_frames * (nseg + 1) * sizeof(int), s);
    if (rc != PSB_OK || n_frames == 0) return rc;
    BinParams p;
    p.pos = positions; p.type_idx = type_idx; p.F = n_frames; p.A = n_atoms; p.nz = nz; p.ntypes = ntypes;
    p.lo = lo; p.hi = hi; p.inv_dz = 1.0 / dz; p.inv_lx = 1.0 / lx_eff; p.inv_ly = 1.0 / ly_eff;
    p.seg_of = seg_scratch; p.offsets = offsets; p.atom_list = atom_list; p.ux = ux; p.uy = uy;
    p.cap = 2 * n_atoms;
    p.unsorted = seg_scratch + (long long)n_frames * n_atoms * 2;
    p.cursor = p.unsorted + (long long)n_frames * p.cap;
    rc = rt::zero(p.cursor, (size_t)n_frames * nseg * sizeof(int), s);
    if (rc != PSB_OK) return rc;
    if (n_atoms > 0) {
        rc = go<BinAssign>(dim3((n_atoms + 255) / 256, n_frames), 0, s, p, "bin_assign");
        if (rc != PSB_OK) return rc;
    }
    rc = go<BinScan>(dim3(n_frames), BinScan::kThreads * sizeof(int), s, p, "bin_scan");
    if (rc != PSB_OK) return rc;
    if (n_atoms > 0) {
        rc = go<BinScatter>(dim3((n_atoms + 255) / 256, n_frames), 0, s, p, "bin_scatter");
        if (rc != PSB_OK) return rc;
        rc = go<BinOrder>(dim3((nseg + 7) / 8, n_frames), 0, s, p, "bin_order");
        if (rc != PSB_OK) return rc;
        // segments too large for one warp's all-pairs ranking (an upper bound per segment is all the host knows)
        if (n_atoms > kBinOrderWarpMax) rc = go<BinOrderLarge>(dim3(nseg, n_frames), BinOrderLarge::kSmem, s, p, "bin_order_large");
    }
    return rc;
}

// phase_out != nullptr: the stack is written as float32 phases sigma*V instead of t = exp(i*sigma*V) (fused grids only)
// `s` is the stream the launches go to (the caller's, or the capture stream while a graph is recorded); `owner` is always
// the caller's stream: the pipelined structure factor keeps its workspace per (device, caller stream)
static int build_eager(const int32_t* offsets, const uint32_t* ux, const uint32_t* uy, int n_frames,
                       int n_atoms, int nz, int ntypes, int nx, int ny, const float* formfactors,
                       float scale, float sigma, psb_c64* t_out, float* v_out, float* phase_out, psb_c64* scratch,
                       long long scratch_elems, cudaStream_t s, cudaStream_t owner) {
    if (!offsets || !ux || !uy || !formfactors || (!t_out && !phase_out) || !scratch)
        return fail(PSB_ERR_INVALID, "psb_build_transmission: null pointer");
#ifndef PSB_EMU
    if (phase_out && !fast_slice_supported(nx, ny))
#else
    if (phase_out)
#endif
        return fail(PSB_ERR_UNSUPPORTED, "psb_build_phase: the phase format needs a grid with fused kernels (256 / 512 points)");
    if (n_frames < 0 || nz < 1 || nx < 1 || ny < 1 || ntypes < 1) return fail(PSB_ERR_INVALID, "psb_build_transmission: bad sizes");
    if (n_frames == 0) return PSB_OK;
    const long long img = (long long)nx * ny;
    const int npairs = (nz + 1) / 2;
    // Work in chunks of slice-pair images small enough to stay in L2 between the three kernels
    // (structure factor -> column IFFT -> row IFFT + transmission), so only t itself goes to HBM.
    long long chunk_imgs = scratch_elems / img;
    if (chunk_imgs < 1) return fail(PSB_ERR_INVALID, "psb_build_transmission: scratch smaller than one image (nx*ny elements)");
    if (chunk_imgs > 65535) chunk_imgs = 65535;
    int mc, fc;                                                  // pairs per chunk, frames per chunk
    if (chunk_imgs >= npairs) {
        mc = npairs;
        fc = (int)(chunk_imgs / npairs);
        if (fc > n_frames) fc = n_frames;
    } else {
        mc = (int)chunk_imgs;
        fc = 1;
    }
    const int nseg = nz * ntypes;
    const int TXs = StructureFactorPaired::TX, TYs = StructureFactorPaired::TY;
    const int tiles = ((StructureFactorPaired::slots(nx) + TXs - 1) / TXs) * ((StructureFactorPaired::slots(ny) + TYs - 1) / TYs);
#ifndef PSB_EMU
    // pipelined structure factor (sf_fast.cu) unless the type count exceeds what it stages: then the generic kernel
    const bool sf_nufft = fast_path_enabled() && fast_slice_supported(nx, ny) && sf_nufft_wanted(ntypes, nx, ny, n_atoms, nz);
    const bool sf_fast = !sf_nufft && fast_path_enabled() && sf_fast_supported(ntypes);
    if (sf_fast) {
        int rc0 = sf_fast_prepare(formfactors, ntypes, nx, ny, s, owner);
        if (rc0 != PSB_OK) return rc0;
    }
    if (sf_nufft) {
        int rc0 = sf_nufft_prepare(offsets, ux, uy, 2 * n_atoms, nz, ntypes, nx, ny, n_frames, formfactors, s, owner);
        if (rc0 != PSB_OK) return rc0;
    }
#endif
#ifndef PSB_EMU
    // inverse row transform of a chunk + epilogue: t = exp(i*sigma*V) (and optionally V), or the bare phases
    auto rows_out = [&](int nf, int nm, int f0, int mb) {
        const float norm = scale / ((float)nx * (float)ny);
        if (phase_out)
            return launch_fast_rows_phase_out(f2(scratch), nf * nm, nx, ny, norm, sigma, phase_out + (long long)f0 * nz * img,
                                              nm, nz, mb, s);
        return launch_fast_rows_transmit(f2(scratch), nf * nm, nx, ny, norm, sigma, f2(t_out) + (long long)f0 * nz * img,
                                         v_out ? v_out + (long long)f0 * nz * img : nullptr, nm, nz, mb, s);
    };
#endif
    for (int f0 = 0; f0 < n_frames; f0 += fc) {
        const int nf = n_frames - f0 < fc ? n_frames - f0 : fc;
        for (int mb = 0; mb < npairs; mb += mc) {
            const int nm = npairs - mb < mc ? npairs - mb : mc;
            SfPairParams sp;
            sp.offsets = offsets + (long long)f0 * (nseg + 1);
            sp.ux = ux + (long long)f0 * 2 * n_atoms; sp.uy = uy + (long long)f0 * 2 * n_atoms;
            sp.cap = 2 * n_atoms; sp.nz = nz; sp.ntypes = ntypes; sp.nx = nx; sp.ny = ny;
            sp.pair_begin = mb; sp.pair_count = nm;
            sp.ff = formfactors; sp.out = f2(scratch);
            // several pairs per block only when one pair per block would already overfill the machine
            const long long want_blocks = 8LL * rt::sm_count();
            long long groups = (want_blocks + (long long)tiles * nf - 1) / ((long long)tiles * nf);
            if (groups < 1) groups = 1;
            if (groups > nm) groups = nm;
            sp.pairs_per_block = (int)((nm + groups - 1) / groups);
            groups = (nm + sp.pairs_per_block - 1) / sp.pairs_per_block;
            int rc;
#ifndef PSB_EMU
            if (sf_nufft)
                rc = launch_sf_nufft(offsets, ux, uy, 2 * n_atoms, nz, ntypes, nx, ny, n_frames, f0, nf, mb, nm, formfactors, f2(scratch), s, owner);
            else if (sf_fast)
                rc = launch_sf_fast(sp.offsets, sp.ux, sp.uy, sp.cap, nz, ntypes, nx, ny, mb, nm, nf, f2(scratch), s, owner);
            else
#endif
            rc = go<StructureFactorPaired>(dim3(tiles, (unsigned)groups, nf), StructureFactorPaired::kSmem, s, sp, "structure_factor");
            if (rc != PSB_OK) return rc;

#ifndef PSB_EMU
            if (fast_slice_supported(nx, ny)) {      // persistent TMA kernels (fast_path.cu), same arithmetic
                rc = launch_fast_cols_inverse(f2(scratch), nf * nm, nx, ny, s);
                if (rc != PSB_OK) return rc;
                rc = rows_out(nf, nm, f0, mb);
                if (rc != PSB_OK) return rc;
                continue;
            }
#endif
            PassParams p = base_params();
            p.src = f2(scratch); p.dst = f2(scratch); p.src_img_stride = img; p.dst_img_stride = img;
            cols_geometry(p, nx, ny);
            rc = launch_line_pass(PASS_INV_COLS, p, nf * nm, s);
            if (rc != PSB_OK) return rc;
            rows_geometry(p, nx, ny);
            p.dst = f2(t_out) + (long long)f0 * nz * img;
            p.scale = scale / ((float)nx * (float)ny);
            p.sigma = sigma;
            p.vout = v_out ? v_out + (long long)f0 * nz * img : nullptr;
            p.pair_count = nm; p.pair_nz = nz; p.pair_begin = mb;
            rc = launch_line_pass(PASS_RI2, p, nf * nm, s);
            if (rc != PSB_OK) return rc;
        }
    }
    return PSB_OK;
}

static int build_impl(const int32_t* offsets, const uint32_t* ux, const uint32_t* uy, int n_frames,
                      int n_atoms, int nz, int ntypes, int nx, int ny, const float* formfactors,
                      float scale, float sigma, psb_c64* t_out, float* v_out, float* phase_out, psb_c64* scratch,
                      long long scratch_elems, void* stream) {
    cudaStream_t owner = as_stream(stream);
    auto eager = [&](cudaStream_t s) {
        return build_eager(offsets, ux, uy, n_frames, n_atoms, nz, ntypes, nx, ny, formfactors, scale, sigma, t_out, v_out,
                           phase_out, scratch, scratch_elems, s, owner);
    };
#ifndef PSB_EMU
    // same buffers and sizes as an earlier call on this stream -> the recorded launch sequence is replayed as one graph
    struct Key {
        const void *offsets, *ux, *uy, *ff, *t_out, *v_out, *phase_out, *scratch, *owner;
        long long scratch_elems;
        int n_frames, n_atoms, nz, ntypes, nx, ny, level, tag;
        float scale, sigma;
    } key;
    std::memset(&key, 0, sizeof(key));
    key.offsets = offsets; key.ux = ux; key.uy = uy; key.ff = formfactors; key.t_out = t_out; key.v_out = v_out;
    key.phase_out = phase_out; key.scratch = scratch; key.owner = owner; key.scratch_elems = scratch_elems;
    key.n_frames = n_frames; key.n_atoms = n_atoms; key.nz = nz; key.ntypes = ntypes; key.nx = nx; key.ny = ny;
    key.level = fast_path_level() * 8 + sf_mode(); key.tag = 0x6275696c;
    key.scale = scale; key.sigma = sigma;
    return run_graphed(&key, sizeof(key), owner, eager);
#else
    return eager(owner);
#endif
}

int psb_build_transmission(const int32_t* offsets, const uint32_t* ux, const uint32_t* uy, int n_frames,
                           int n_atoms, int nz, int ntypes, int nx, int ny, const float* formfactors,
                           float scale, float sigma, psb_c64* t_out, float* v_out, psb_c64* scratch,
                           long long scratch_elems, void* stream) {
    if (!t_out) return fail(PSB_ERR_INVALID, "psb_build_transmission: null pointer");
    return build_impl(offsets, ux, uy, n_frames, n_atoms, nz, ntypes, nx, ny, formfactors, scale, sigma, t_out, v_out, nullptr,
                      scratch, scratch_elems, stream);
}

int psb_build_phase(const int32_t* offsets, const uint32_t* ux, const uint32_t* uy, int n_frames,
                    int n_atoms, int nz, int ntypes, int nx, int ny, const float* formfactors,
                    float scale, float sigma, float* phase_out, psb_c64* scratch, long long scratch_elems, void* stream) {
    if (!phase_out) return fail(PSB_ERR_INVALID, "psb_build_phase: null pointer");
    return build_impl(offsets, ux, uy, n_frames, n_atoms, nz, ntypes, nx, ny, formfactors, scale, sigma, nullptr, nullptr,
                      phase_out, scratch, scratch_elems, stream);
}

int psb_phase_format_supported(int nx, int ny) {
#ifndef PSB_EMU
    return fast_slice_supported(nx, ny) ? 1 : 0;
#else
    (void)nx; (void)ny;
    return 0;
#endif
}

int psb_transmission_from_potential(const float* v, psb_c64* t, long long n, float sigma, void* stream) {
    if (!v || !t || n < 0) return fail(PSB_ERR_INVALID, "psb_transmission_from_potential: bad argument");
    if (n == 0) return PSB_OK;
    TransmitParams p{v, f2(t), n, sigma, n, n};
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    return go<Transmit>(dim3((unsigned)blocks), 0, as_stream(stream), p, "transmit");
}

int psb_fft2(const psb_c64* src, psb_c64* dst, int batch, int nx, int ny, int inverse, float scale, void* stream) {
    if (!src || !dst || batch < 0 || nx < 1 || ny < 1) return fail(PSB_ERR_INVALID, "psb_fft2: bad argument");
    cudaStream_t s = as_stream(stream);
    const long long img = (long long)nx * ny;
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
        PassParams p = base_params();
        p.src = f2(src) + b0 * img; p.dst = f2(dst) + b0 * img; p.src_img_stride = img; p.dst_img_stride = img;
        rows_geometry(p, nx, ny);
        int rc = launch_line_pass(inverse ? PASS_INV_ROWS : PASS_FWD_ROWS, p, nb, s);
        if (rc != PSB_OK) return rc;
        p.src = p.dst;
        cols_geometry(p, nx, ny);
        p.scale = scale;
        rc = launch_line_pass(inverse ? PASS_INV_COLS : PASS_FWD_COLS, p, nb, s);
        if (rc != PSB_OK) return rc;
    }
    return PSB_OK;
}

int psb_shift_probes(const psb_c64* base_k, const psb_c64* ramp_x, const psb_c64* ramp_y, int n_probes,
                     int nx, int ny, psb_c64* out, void* stream) {
    if (!base_k || !ramp_x || !ramp_y || !out || n_probes < 0 || nx < 1 || ny < 1)
        return fail(PSB_ERR_INVALID, "psb_shift_probes: bad argument");
    if (n_probes > 65535) return fail(PSB_ERR_UNSUPPORTED, "psb_shift_probes: more than 65535 probes");
    cudaStream_t s = as_stream(stream);
    const long long img = (long long)nx * ny;
    PassParams p = base_params();
    p.src = f2(base_k); p.dst = f2(out); p.src_img_stride = 0; p.dst_img_stride = img;
    cols_geometry(p, nx, ny);
    p.sep_p = f2(ramp_x); p.sep_img_stride_p = nx;    // along the line (kx)
    p.sep_l = f2(ramp_y); p.sep_img_stride_l = ny;    // per line (ky)
    int rc = launch_line_pass(PASS_CP, p, n_probes, s);
    if (rc != PSB_OK) return rc;
    p = base_params();
    p.src = f2(out); p.dst = f2(out); p.src_img_stride = img; p.dst_img_stride = img;
    rows_geometry(p, nx, ny);
    p.scale = 1.0f / ((float)nx * (float)ny);
    return launch_line_pass(PASS_INV_ROWS, p, n_probes, s);
}

static int sum_pixels_impl(const float* in, const float2* cin, const float* mask, int rows, long long row_stride,
                           long long npix, double* out, void* stream, int row_mod = 0, long long out_stride_mod = 0,
                           long long out_stride_div = 0);

// the launch sequence of one psb_propagate_ex call (arguments already validated), issued on stream `s`
static int propagate_eager(const psb_propagate_desc& d, cudaStream_t s) {
    const psb_c64* t = d.t;
    const float* phase = d.phase;
    const int n_frames = d.n_frames, n_probes = d.n_probes, nz = d.nz, nx = d.nx, ny = d.ny, mode = d.mode;
    const bool slabs = mode == 1 && d.slab_world > 1;
    const long long img = (long long)nx * ny;
    const int n_img = n_frames * n_probes;
    const int layer_every = d.layer_every;
    psb_c64* psi_work = d.psi_work;

    PassParams row = base_params();
    row.dst = f2(psi_work); row.dst_img_stride = img;
    rows_geometry(row, nx, ny);
    row.mul_img_stride = (long long)nz * img;
    row.mul_img_div = n_probes;

    PassParams col = base_params();
    col.src = f2(psi_work); col.dst = f2(psi_work); col.src_img_stride = img; col.dst_img_stride = img;
    cols_geometry(col, nx, ny);
    col.sep_p = f2(d.prop_x); col.sep_l = f2(d.prop_y);

    PassParams ex = base_params();
    ex.src = f2(psi_work); ex.src_img_stride = img;
    cols_geometry(ex, nx, ny);
    ex.probes = n_probes;
    ex.out_stride_probe = d.stride_probe; ex.out_stride_frame = d.stride_frame;
    ex.out_elem_stride = ny; ex.out_line_stride = 1;
    if (slabs) {
        ex.slab_world = d.slab_world; ex.slab_base = nx / d.slab_world; ex.slab_rem = nx % d.slab_world;
        ex.slab_planes = (long long)d.slab_layers * d.slab_frames * d.slab_probes;
        ex.out_stride_probe = 1; ex.out_stride_frame = d.slab_probes;          // in planes
    }
    if (mode == 2) {                                                           // dense (frame, probe) images in det_scratch
        ex.out_stride_probe = img; ex.out_stride_frame = (long long)n_probes * img;
    }

#ifndef PSB_EMU
    // fused persistent kernels for the steady state (fast_path.cu)
    const bool fast = fast_slice_supported(nx, ny);
#else
    const bool fast = false;
#endif
    if (phase) {
        if (!fast) return fail(PSB_ERR_UNSUPPORTED, "psb_propagate_phase: the phase format needs a grid with fused kernels");
        // slice 0 of every frame as t for the generic first pass (one strided launch)
        TransmitParams tp{phase, f2(d.t0_scratch), (long long)n_frames * img, 1.0f, img, (long long)nz * img};
        long long blocks = (tp.n + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        int rc0 = go<Transmit>(dim3((unsigned)blocks), 0, s, tp, "transmit");
        if (rc0 != PSB_OK) return rc0;
        row.mul_img_stride = img;
    }
    int layer = 0;
    for (int z = 0; z < nz; ++z) {
        row.mul = phase ? f2(d.t0_scratch) : f2(t) + (long long)z * img;
        int rc;
        if (z > 0 && fast) {
#ifndef PSB_EMU
            if (phase)
                rc = launch_fast_rows_phase(f2(psi_work), n_img, nx, ny, phase + (long long)z * img, (long long)nz * img, n_probes, s);
            else
                rc = launch_fast_rows(f2(psi_work), n_img, nx, ny, f2(t) + (long long)z * img, (long long)nz * img, n_probes, s);
#endif
        } else if (z == 0) {
            row.src = f2(d.probes); row.src_img_stride = img; row.src_img_mod = n_probes;
            rc = launch_line_pass(PASS_R1, row, n_img, s);
        } else {
            row.src = f2(psi_work); row.src_img_stride = img; row.src_img_mod = 0;
            rc = launch_line_pass(PASS_R, row, n_img, s);
        }
        if (rc != PSB_OK) return rc;
        const bool last = z == nz - 1;
        const bool tap = mode >= 1 && (last || (layer_every > 0 && (z + 1) % layer_every == 0));
        if (tap) {
            if (mode == 2) {
                ex.dst = f2(d.det_scratch);
            } else if (slabs) {
                ex.dst = f2(d.wf_out);
                ex.slab_plane0 = ((long long)layer * d.slab_frames + d.frame0) * d.slab_probes + d.probe0;
            } else {
                ex.dst = f2(d.wf_out) + (long long)layer * d.stride_layer;
            }
            rc = launch_line_pass(PASS_CX, ex, n_img, s);
            if (rc != PSB_OK) return rc;
            if (mode == 2) {          // image index = frame*n_probes + probe -> det_out[layer][probe0 + probe][frame0 + frame]
                rc = sum_pixels_impl(nullptr, f2(d.det_scratch), d.det_mask, n_img, img, img,
                                     d.det_out + (long long)layer * d.det_stride_layer + (long long)d.probe0 * d.det_stride_probe + d.frame0,
                                     (void*)s, n_probes, d.det_stride_probe, 1);
                if (rc != PSB_OK) return rc;
            }
            ++layer;
        }
        if (!last) {
#ifndef PSB_EMU
            if (fast) rc = launch_fast_cols(f2(psi_work), n_img, nx, ny, f2(d.prop_x), f2(d.prop_y), s);
            else
#endif
            rc = launch_line_pass(PASS_C, col, n_img, s);
            if (rc != PSB_OK) return rc;
        }
    }
    if (mode == 0) {   // back to real space along y: psi is [x, ky] after the last row pass
        PassParams inv = base_params();
        inv.src = f2(psi_work); inv.dst = f2(psi_work); inv.src_img_stride = img; inv.dst_img_stride = img;
        rows_geometry(inv, nx, ny);
        inv.scale = 1.0f / (float)ny;
        return launch_line_pass(PASS_INV_ROWS, inv, n_img, s);
    }
    return PSB_OK;
}

// desc.phase != nullptr: the stack holds float32 phases; t0_scratch (n_frames, nx, ny) receives exp(i*phase) of slice 0 for
// the generic first pass, every later slice is evaluated inside the fused row pass
int psb_propagate_ex(const psb_propagate_desc* dp) {
    if (!dp || dp->struct_bytes != (int)sizeof(psb_propagate_desc))
        return fail(PSB_ERR_INVALID, "psb_propagate_ex: descriptor size mismatch (header / library version skew)");
    const psb_propagate_desc& d = *dp;
    const psb_c64* t = d.t;
    const float* phase = d.phase;
    const int n_frames = d.n_frames, n_probes = d.n_probes, nz = d.nz, nx = d.nx, ny = d.ny, mode = d.mode;
    if (!d.probes || (!t && !phase) || (t && phase) || !d.prop_x || !d.prop_y || !d.psi_work)
        return fail(PSB_ERR_INVALID, "psb_propagate: null pointer (or both t and phase given)");
    if (phase && !d.t0_scratch) return fail(PSB_ERR_INVALID, "psb_propagate_phase: t0 scratch required");
    if (mode < 0 || mode > 2) return fail(PSB_ERR_INVALID, "psb_propagate: mode must be 0, 1 or 2");
    if (mode == 1 && !d.wf_out) return fail(PSB_ERR_INVALID, "psb_propagate: wf_out required in mode 1");
    if (mode == 2 && (!d.det_mask || !d.det_out || !d.det_scratch))
        return fail(PSB_ERR_INVALID, "psb_propagate: det_mask, det_out and det_scratch required in mode 2");
    if (n_frames < 0 || n_probes < 1 || nz < 1 || nx < 1 || ny < 1 || d.layer_every < 0)
        return fail(PSB_ERR_INVALID, "psb_propagate: bad sizes");
    const bool slabs = mode == 1 && d.slab_world > 1;
    if (slabs && (d.slab_world > nx || d.slab_layers < 1 || d.slab_frames < 1 || d.slab_probes < 1 || d.frame0 < 0 || d.probe0 < 0 ||
                  d.frame0 + n_frames > d.slab_frames || d.probe0 + n_probes > d.slab_probes))
        return fail(PSB_ERR_INVALID, "psb_propagate: bad slab layout");
    const long long n_img_ll = (long long)n_frames * n_probes;
    if (n_img_ll == 0) return PSB_OK;
    cudaStream_t s = as_stream(d.stream);
    const long long img = (long long)nx * ny;
    // the generic passes put the image index in gridDim.y: larger batches go through in slices of whole frames
    if (n_img_ll > 65535) {
        const int fmax = 65535 / n_probes;
        if (fmax < 1) return fail(PSB_ERR_UNSUPPORTED, "psb_propagate: more than 65535 probes per call");
        for (int f0 = 0; f0 < n_frames; f0 += fmax) {
            psb_propagate_desc part = d;
            part.n_frames = n_frames - f0 < fmax ? n_frames - f0 : fmax;
            if (t) part.t = t + (long long)f0 * nz * img;
            if (phase) part.phase = phase + (long long)f0 * nz * img;
            if (mode == 1 && !slabs) part.wf_out = d.wf_out + (long long)f0 * d.stride_frame;
            part.psi_work = d.psi_work + (long long)f0 * n_probes * img;          // mode 0 leaves its result there
            part.frame0 = d.frame0 + f0;
            int rc = psb_propagate_ex(&part);
            if (rc != PSB_OK) return rc;
        }
        return PSB_OK;
    }
    // same buffers and sizes as an earlier call -> the recorded launch sequence (two kernels per slice) replays as one graph
    auto eager = [&](cudaStream_t ls) { return propagate_eager(d, ls); };
#ifndef PSB_EMU
    struct Key {
        psb_propagate_desc d;
        int level, tag;
    } key;
    std::memset(&key, 0, sizeof(key));
    std::memcpy(&key.d, &d, sizeof(d));          // includes padding bytes of d, which callers zero (memset / ctypes)
    key.d.stream = nullptr;
    key.level = fast_path_level(); key.tag = 0x70726f70;
    return run_graphed(&key, sizeof(key), s, eager);
#else
    return eager(s);
#endif
}

static psb_propagate_desc dense_desc(const psb_c64* probes, int n_frames, int n_probes, int nz, int nx, int ny, const psb_c64* prop_x,
                                     const psb_c64* prop_y, psb_c64* psi_work, int mode, psb_c64* wf_out, long long stride_probe,
                                     long long stride_frame, long long stride_layer, int layer_every, void* stream) {
    psb_propagate_desc d;
    std::memset(&d, 0, sizeof(d));
    d.struct_bytes = (int)sizeof(d);
    d.mode = mode; d.layer_every = layer_every;
    d.n_frames = n_frames; d.n_probes = n_probes; d.nz = nz; d.nx = nx; d.ny = ny;
    d.probes = probes; d.prop_x = prop_x; d.prop_y = prop_y; d.psi_work = psi_work; d.wf_out = wf_out;
    d.stride_probe = stride_probe; d.stride_frame = stride_frame; d.stride_layer = stride_layer;
    d.stream = stream;
    return d;
}

int psb_propagate(const psb_c64* probes, const psb_c64* t, int n_frames, int n_probes, int nz, int nx, int ny,
                  const psb_c64* prop_x, const psb_c64* prop_y, psb_c64* psi_work, int mode, psb_c64* wf_out,
                  long long stride_probe, long long stride_frame, long long stride_layer, int layer_every,
                  void* stream) {
    if (!t) return fail(PSB_ERR_INVALID, "psb_propagate: null pointer");
    if (mode != 0 && mode != 1) return fail(PSB_ERR_INVALID, "psb_propagate: mode must be 0 or 1");
    psb_propagate_desc d = dense_desc(probes, n_frames, n_probes, nz, nx, ny, prop_x, prop_y, psi_work, mode, wf_out, stride_probe,
                                      stride_frame, stride_layer, layer_every, stream);
    d.t = t;
    return psb_propagate_ex(&d);
}

int psb_propagate_phase(const psb_c64* probes, const float* phase, psb_c64* t0, int n_frames, int n_probes, int nz, int nx,
                        int ny, const psb_c64* prop_x, const psb_c64* prop_y, psb_c64* psi_work, int mode, psb_c64* wf_out,
                        long long stride_probe, long long stride_frame, long long stride_layer, int layer_every,
                        void* stream) {
    if (!phase) return fail(PSB_ERR_INVALID, "psb_propagate_phase: null pointer");
    if (mode != 0 && mode != 1) return fail(PSB_ERR_INVALID, "psb_propagate: mode must be 0 or 1");
    psb_propagate_desc d = dense_desc(probes, n_frames, n_probes, nz, nx, ny, prop_x, prop_y, psi_work, mode, wf_out, stride_probe,
                                      stride_frame, stride_layer, layer_every, stream);
    d.phase = phase; d.t0_scratch = t0;
    return psb_propagate_ex(&d);
}

int psb_tacaw_intensity(const psb_c64* wf, long long stride_probe, long long stride_frame, int n_probes,
                        int n_frames, long long npix, float* intensity, void* stream) {
    if (!wf || !intensity || n_probes < 0 || n_frames < 1 || npix < 0) return fail(PSB_ERR_INVALID, "psb_tacaw_intensity: bad argument");
    if (npix > 0x7fffffffLL) return fail(PSB_ERR_UNSUPPORTED, "psb_tacaw_intensity: npix too large");
    if (n_probes > 65535) return fail(PSB_ERR_UNSUPPORTED, "psb_tacaw_intensity: more than 65535 probes");
#ifndef PSB_EMU
    // frame counts 2^a 3^b 5^c: tiled mixed-radix transform (tacaw_fast.cu); anything else: Bluestein line pass
    if (fast_path_enabled() && tacaw_fast_supported(n_frames))
        return launch_tacaw_fast(f2(wf), stride_probe, stride_frame, n_probes, n_frames, npix, intensity, as_stream(stream));
#endif
    PassParams p = base_params();
    p.src = f2(wf); p.src_img_stride = stride_probe;
    p.nlines = (int)npix; p.line_len = n_frames; p.line_stride = 1; p.elem_stride = stride_frame;
    p.fout = intensity; p.dst_img_stride = (long long)n_frames * npix;
    p.out_elem_stride = npix; p.out_line_stride = 1;
    return launch_line_pass(PASS_TW, p, n_probes, as_stream(stream));
}

static int sum_pixels_impl(const float* in, const float2* cin, const float* mask, int rows, long long row_stride,
                           long long npix, double* out, void* stream, int row_mod, long long out_stride_mod,
                           long long out_stride_div) {
    if ((!in && !cin) || !out || rows < 0 || npix < 0) return fail(PSB_ERR_INVALID, "psb_sum_pixels: bad argument");
    if (rows == 0) return PSB_OK;
    SumPixParams p{in, cin, mask, row_stride, npix, out, row_mod, out_stride_mod, out_stride_div};
    return go<SumPixels>(dim3(rows), SumPixels::kSmem, as_stream(stream), p, "sum_pixels");
}

int psb_sum_pixels(const float* in, const float* mask, int rows, long long row_stride, long long npix,
                   double* out, void* stream) {
    return sum_pixels_impl(in, nullptr, mask, rows, row_stride, npix, out, stream);
}

int psb_sum_abs_pixels(const psb_c64* in, const float* mask, int rows, long long row_stride, long long npix,
                       double* out, void* stream) {
    return sum_pixels_impl(nullptr, f2(in), mask, rows, row_stride, npix, out, stream);
}

int psb_sum_frames(const float* in, int groups, int n_frames, long long npix, float* out, void* stream) {
    if (!in || !out || groups < 0 || n_frames < 1 || npix < 0) return fail(PSB_ERR_INVALID, "psb_sum_frames: bad argument");
    if (groups == 0 || npix == 0) return PSB_OK;
    if (groups > 65535) return fail(PSB_ERR_UNSUPPORTED, "psb_sum_frames: more than 65535 groups");
    SumRowsParams p{in, n_frames, npix, out};
    return go<SumRows>(dim3((unsigned)((npix + 255) / 256), groups), 0, as_stream(stream), p, "sum_frames");
}

}  // extern "C"
