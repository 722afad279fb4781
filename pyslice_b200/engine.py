"""Host-side driver of the CUDA engine: owns device buffers (torch tensors used purely as
containers), frame-invariant tables, and the batching of frames / probes through libpsb.

Nothing here computes on the CPU beyond the once-per-setup float64 tables of `hostmath`; every
per-frame operation is a libpsb call on CUDA memory.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, hostmath
from ctypes import sizeof as C_sizeof

# Target size of the wave-function batch that is pushed through all slices at once.  Sized to stay
# resident in B200's 126 MB L2 across the two passes of a slice step (DESIGN.md "L2 residency"): the
# transmission stack streams through L2 with an evict-first hint, so ~3/5 of L2 can hold psi
# (measured: 148 images of 256^2 = 74 MB still resident, 185 = 93 MB not; profiles/r1e notes).
PSI_BATCH_BYTES = 80 << 20
# 1024-point grids: one image is 8 MB and a batch that fills the persistent grids never fits L2, so psi streams from HBM
# whatever the batch; then larger batches only amortise launch ramps and tails (measured 9 / 16 / 24 / 37 images:
# 110 / 112 / 115 / 119 k slice-steps/s, profiles/r2ad_batch_sweep.txt; 256 and 512 points lose 13-18 % past L2).
PSI_BATCH_BYTES_STREAMING = 320 << 20
# Upper bound for the transmission-function buffer of one frame batch.
T_BATCH_BYTES = 48 << 30
# Workspace of the potential build (slice-paired spectra): chunks of this size flow through the three
# potential kernels while staying L2-resident.
SCRATCH_BYTES = 64 << 20


# per-device batch workspaces of the calculators (multislice/calculators.py: run()), kept so that consecutive runs and
# consecutive calculators of one geometry see the same device addresses (CUDA-graph replay in libpsb keys on them)
WORKSPACES = {}
_FF_CACHE = {}      # (device, grid, types) -> form-factor table on the device (make_plan)


class PhaseTimer:
    """Optional CUDA-event timing of the pipeline phases on the launching stream (used by bench.py to
    report the slice-step kernel's own duration).  `with timer.phase("propagate"): ...` records an
    event pair; `totals()` synchronises and returns milliseconds per phase."""

    def __init__(self, device):
        self.device = device
        self.pairs = []

    class _Span:
        def __init__(self, owner, name):
            self.owner, self.name = owner, name

        def __enter__(self):
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record(torch.cuda.current_stream(self.owner.device))

        def __exit__(self, *exc):
            self.b.record(torch.cuda.current_stream(self.owner.device))
            self.owner.pairs.append((self.name, self.a, self.b))

    def phase(self, name):
        return PhaseTimer._Span(self, name)

    def totals(self):
        torch.cuda.synchronize(self.device)
        out = {}
        for name, a, b in self.pairs:
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        self.pairs = []
        return out


class _NoTimer:
    class _Null:
        def __enter__(self):
            return None

        def __exit__(self, *exc):
            return False

    def phase(self, name):
        return _NoTimer._Null()


NO_TIMER = _NoTimer()


def launch_count() -> int:
    return int(_lib.lib().psb_launch_count())


def set_fast_path(enable) -> None:
    """Diagnostic switch between the kernel generations (all CUDA; used by A/B parity tests and microbenchmarks):
    True / 1 = fused persistent kernels (default), False / 0 = generic line-pass kernels."""
    level = 1 if enable is True else (0 if enable is False else int(enable))
    _lib.lib().psb_set_fast_path(level)


def set_sf_mode(mode: int) -> None:
    """Structure-factor algorithm of the potential build: 0 automatic (default), 1 direct sum always, 2 the 1-D NUFFT of
    csrc/sf_nufft.cu wherever it is implemented (A/B tests and microbenchmarks)."""
    _lib.lib().psb_set_sf_mode(int(mode))


def set_graph_mode(on: bool) -> None:
    """CUDA-graph replay of repeated launch sequences inside libpsb (default on); off = every kernel launched directly."""
    _lib.lib().psb_set_graph_mode(1 if on else 0)


def _device(device=None) -> torch.device:
    if _lib.is_emulated():                       # tests/emu only
        return torch.device("cpu")
    if not torch.cuda.is_available():
        raise RuntimeError("pyslice_b200 requires a CUDA device (built for sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"pyslice_b200 runs on CUDA devices only, got {device}")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _stream(device: torch.device) -> int:
    if device.type != "cuda":
        return 0
    return torch.cuda.current_stream(device).cuda_stream


class _on:
    """`with _on(device):` makes `device` the current CUDA device around a libpsb call.  The library launches on the
    CURRENT device and keys its tables and workspaces by it, while every buffer it is handed lives on the tensor's
    device: with several GPUs in one process (MultisliceCalculator(device='cuda:1') while cuda:0 is current) the two
    must be made to agree here.  No-op on the kernel emulator (tests) and when the device is already current."""

    def __init__(self, device):
        self.guard = None
        if device is not None and getattr(device, "type", None) == "cuda" and not _lib.is_emulated():
            idx = device.index if device.index is not None else torch.cuda.current_device()
            if idx != torch.cuda.current_device():
                self.guard = torch.cuda.device(idx)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()
        return self

    def __exit__(self, *exc):
        if self.guard is not None:
            return self.guard.__exit__(*exc)
        return False


def _c64(x: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(x).astype(np.complex64)).to(device)


@dataclass
class SlicePlan:
    """Everything that does not change from frame to frame for one (box, sampling, voltage, types)."""
    device: torch.device
    xs: np.ndarray
    ys: np.ndarray
    zs: np.ndarray
    nx: int
    ny: int
    nz: int
    dx: float
    dy: float
    dz: float
    eV: float
    wavelength: float
    sigma: float
    type_Z: list                  # sorted unique atomic numbers
    type_idx: torch.Tensor        # (A,) int32 on device
    lo: torch.Tensor              # (nz,) float64 on device
    hi: torch.Tensor
    formfactors: torch.Tensor     # (ntypes, nx, ny) float32
    prop_x: torch.Tensor          # (nx,) complex64, includes 1/(nx*ny)
    prop_y: torch.Tensor          # (ny,) complex64
    kxs: np.ndarray
    kys: np.ndarray

    @property
    def ntypes(self) -> int:
        return len(self.type_Z)


def make_plan(xs, ys, zs, atom_kinds: Sequence, eV: float, device=None) -> SlicePlan:
    device = _device(device)
    xs = np.asarray(xs, dtype=np.float64)
    ys = np.asarray(ys, dtype=np.float64)
    zs = np.asarray(zs, dtype=np.float64)
    nx, ny, nz = len(xs), len(ys), len(zs)
    Z = np.array([hostmath.atomic_number(k) for k in atom_kinds], dtype=np.int64)
    type_Z = sorted(set(Z.tolist()))
    lut = {z: i for i, z in enumerate(type_Z)}
    type_idx = np.array([lut[z] for z in Z.tolist()], dtype=np.int32)
    lo, hi, dz = hostmath.slice_bounds(zs)
    kxs, kys = hostmath.kgrid(xs, ys)
    # Kirkland form factors on the k grid: a pure function of (grid, atom types), 6-40 ms of NumPy per setup() at 256^2 ...
    # 1024^2 -- kept per device for the next calculator of the same geometry (a few entries, oldest dropped)
    ff_key = (str(device), nx, ny, float(xs[1] - xs[0]), float(ys[1] - ys[0]), tuple(type_Z))
    ff_dev = _FF_CACHE.get(ff_key)
    if ff_dev is None:
        ff_dev = torch.from_numpy(hostmath.form_factor_table(kxs, kys, type_Z).astype(np.float32)).to(device)
        if len(_FF_CACHE) >= 4:
            _FF_CACHE.pop(next(iter(_FF_CACHE)))
        _FF_CACHE[ff_key] = ff_dev
    lam = hostmath.wavelength(eV)
    px, py = hostmath.propagator_tables(kxs, kys, lam, dz)
    px = px / (nx * ny)
    return SlicePlan(
        device=device, xs=xs, ys=ys, zs=zs, nx=nx, ny=ny, nz=nz, dx=float(xs[1] - xs[0]), dy=float(ys[1] - ys[0]),
        dz=dz, eV=eV, wavelength=lam, sigma=hostmath.interaction_sigma(eV), type_Z=type_Z,
        type_idx=torch.from_numpy(type_idx).to(device), lo=torch.from_numpy(lo).to(device),
        hi=torch.from_numpy(hi).to(device), formfactors=ff_dev,
        prop_x=_c64(px, device), prop_y=_c64(py, device), kxs=kxs, kys=kys)


# --------------------------------------------------------------------------------------------
def fft2(x: torch.Tensor, inverse: bool = False, scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batched 2-D FFT of a (..., nx, ny) complex64 tensor through psb_fft2 (unnormalised; `scale`
    multiplies the result, pass 1/(nx*ny) for torch's ifft2 convention)."""
    assert x.dtype == torch.complex64 and x.is_contiguous()
    nx, ny = x.shape[-2:]
    batch = x.numel() // (nx * ny)
    if out is None:
        out = torch.empty_like(x)
    L = _lib.lib()
    with _on(x.device):
        _lib.check(L.psb_fft2(x.data_ptr(), out.data_ptr(), batch, nx, ny, 1 if inverse else 0, scale, _stream(x.device)), "psb_fft2")
    return out


def bin_buffers(plan: SlicePlan, F: int, A: int):
    """scratch + outputs of psb_bin_atoms for up to F frames of A atoms (allocate once per run: stable addresses let
    libpsb replay the potential build of later batches as a recorded graph)"""
    dev = plan.device
    nseg = plan.nz * plan.ntypes
    return dict(F=F, A=A,
                seg=torch.empty((F * (4 * A + nseg),), dtype=torch.int32, device=dev),     # seg ids + unsorted lists + cursors
                offsets=torch.empty((F, nseg + 1), dtype=torch.int32, device=dev),
                atom_list=torch.empty((F, 2 * A), dtype=torch.int32, device=dev),
                ux=torch.empty((F, 2 * A), dtype=torch.int32, device=dev),
                uy=torch.empty((F, 2 * A), dtype=torch.int32, device=dev))


def bin_atoms(plan: SlicePlan, positions: torch.Tensor, buffers=None):
    """positions (F, A, 3) float64 on device -> (offsets (F, nseg+1), atom_list, ux, uy (F, 2A))."""
    F, A, _ = positions.shape
    dev = plan.device
    if buffers is None or buffers["F"] < F or buffers["A"] != A:
        buffers = bin_buffers(plan, F, A)
    seg = buffers["seg"]
    offsets, atom_list, ux, uy = (buffers[k][:F] for k in ("offsets", "atom_list", "ux", "uy"))
    L = _lib.lib()
    with _on(dev):
        _lib.check(L.psb_bin_atoms(positions.data_ptr(), plan.type_idx.data_ptr(), F, A, plan.ntypes, plan.nz,
                                   plan.lo.data_ptr(), plan.hi.data_ptr(), plan.dz, plan.nx * plan.dx, plan.ny * plan.dy,
                                   seg.data_ptr(), offsets.data_ptr(), atom_list.data_ptr(), ux.data_ptr(), uy.data_ptr(),
                                   _stream(dev)), "psb_bin_atoms")
    return offsets, atom_list, ux, uy


def chunk_images(plan: SlicePlan, n_frames: int) -> int:
    """Slice-pair images per chunk of the potential build: as many as SCRATCH_BYTES holds.  For large grids, where the
    structure factor dominates and one image is many tiles (1024 x 1024: 128 tiles of 64 x 32 slots), the count is
    nudged (up to +25 %) to the one with the fewest rounds of the 2-CTAs-per-SM grid per image: 9 images = 1152 tiles
    = 3.9 rounds on 148 SMs where 8 images cost 4 rounds for 3.5 rounds of work.  (At 256 x 256 the same rule -- 111
    instead of 128 images -- was measured and lost: the third, small chunk per frame costs more than the partial round.)"""
    img = plan.nx * plan.ny
    want = max(1, min(n_frames * ((plan.nz + 1) // 2), SCRATCH_BYTES // (8 * img)))
    tiles = -(-(plan.nx // 2) // 64) * -(-(plan.ny // 2) // 32)
    if tiles >= 64 and want > 1:
        slots = 2 * _sm_count()
        best, best_cost = want, None
        for n in range(max(1, want - want // 4), want + want // 4 + 1):
            cost = -(-n * tiles // slots) / n
            if best_cost is None or cost < best_cost - 1e-12:
                best, best_cost = n, cost
        want = min(best, n_frames * ((plan.nz + 1) // 2))
    return want


def phase_format_supported(plan: SlicePlan) -> bool:
    """True when the transmission stack of this grid may be kept as float32 phases (fused kernels: 256 / 512 / 1024 points)."""
    return bool(_lib.lib().psb_phase_format_supported(plan.nx, plan.ny))


def build_transmission(plan: SlicePlan, positions: torch.Tensor, want_potential: bool = False,
                       out: Optional[torch.Tensor] = None, scratch: Optional[torch.Tensor] = None, phase: bool = False,
                       bins=None):
    """positions (F, A, 3) float64 on device -> t (F, nz, nx, ny) complex64 [, V float32].
    phase=True: the stack as float32 phases sigma*V instead (half the bytes; `propagate` accepts either).
    bins: optional `bin_buffers` to reuse across calls."""
    F, A, _ = positions.shape
    dev = plan.device
    offsets, _, ux, uy = bin_atoms(plan, positions, bins)
    scale = 1.0 / (plan.dx ** 2 * plan.dy ** 2)
    img = plan.nx * plan.ny
    n_scratch = img * chunk_images(plan, F)
    if scratch is None or scratch.numel() < n_scratch:
        scratch = torch.empty((n_scratch,), dtype=torch.complex64, device=dev)
    L = _lib.lib()
    if phase:
        assert not want_potential
        ph = out if out is not None else torch.empty((F, plan.nz, plan.nx, plan.ny), dtype=torch.float32, device=dev)
        assert ph.dtype == torch.float32
        with _on(dev):
            _lib.check(L.psb_build_phase(offsets.data_ptr(), ux.data_ptr(), uy.data_ptr(), F, A, plan.nz, plan.ntypes,
                                         plan.nx, plan.ny, plan.formfactors.data_ptr(), scale, plan.sigma,
                                         ph.data_ptr(), scratch.data_ptr(), n_scratch, _stream(dev)), "psb_build_phase")
        return ph
    t = out if out is not None else torch.empty((F, plan.nz, plan.nx, plan.ny), dtype=torch.complex64, device=dev)
    V = torch.empty((F, plan.nz, plan.nx, plan.ny), dtype=torch.float32, device=dev) if want_potential else None
    with _on(dev):
        _lib.check(L.psb_build_transmission(offsets.data_ptr(), ux.data_ptr(), uy.data_ptr(), F, A, plan.nz, plan.ntypes,
                                            plan.nx, plan.ny, plan.formfactors.data_ptr(), scale, plan.sigma,
                                            t.data_ptr(), V.data_ptr() if V is not None else None, scratch.data_ptr(),
                                            n_scratch, _stream(dev)),
                   "psb_build_transmission")
    return (t, V) if want_potential else t


def transmission_from_potential(V: torch.Tensor, sigma: float) -> torch.Tensor:
    assert V.dtype == torch.float32 and V.is_contiguous()
    t = torch.empty(V.shape, dtype=torch.complex64, device=V.device)
    with _on(V.device):
        _lib.check(_lib.lib().psb_transmission_from_potential(V.data_ptr(), t.data_ptr(), V.numel(), sigma, _stream(V.device)),
                   "psb_transmission_from_potential")
    return t


def shift_probes(base_k: torch.Tensor, ramp_x: torch.Tensor, ramp_y: torch.Tensor) -> torch.Tensor:
    """base_k (nx, ny), ramps (P, nx) / (P, ny) complex64 -> (P, nx, ny) shifted probes."""
    P = ramp_x.shape[0]
    nx, ny = base_k.shape
    out = torch.empty((P, nx, ny), dtype=torch.complex64, device=base_k.device)
    with _on(base_k.device):
        _lib.check(_lib.lib().psb_shift_probes(base_k.data_ptr(), ramp_x.data_ptr(), ramp_y.data_ptr(), P, nx, ny,
                                               out.data_ptr(), _stream(base_k.device)), "psb_shift_probes")
    return out


def layer_count(nz: int, layer_every: int) -> int:
    if layer_every <= 0:
        return 1
    return len([z for z in range(nz - 1) if (z + 1) % layer_every == 0]) + 1


def row_split(n: int, world: int):
    """kx rows per destination of the frames -> rows all-to-all: (starts, counts), the first n % world blocks one row longer
    (the rule psb_propagate_ex applies to its slab layout)"""
    counts = [n // world + (1 if r < n % world else 0) for r in range(world)]
    starts = [sum(counts[:r]) for r in range(world)]
    return starts, counts


def propagate(plan: SlicePlan, probes: torch.Tensor, t: torch.Tensor, wf_out: Optional[torch.Tensor] = None,
              frame0: int = 0, probe0: int = 0, layer_every: int = 0, work: Optional[torch.Tensor] = None,
              slabs=None, detector=None, t0: Optional[torch.Tensor] = None):
    """Push `probes` (P, nx, ny) through the transmission stack `t` (F, nz, nx, ny; complex64 t or float32 phases).

    wf_out is None: returns the real-space exit waves (F, P, nx, ny)  (the reference's Propagate()).
    wf_out (L, Ptot, Ttot, nx, ny): writes fftshift(fft2(exit)) for frames [frame0, frame0+F) and probes
    [probe0, probe0+P) of every layer (the reference's per-frame worker + copy loop).
    slabs=(world, L, Ttot, Ptot) with a flat wf_out of L*Ptot*Ttot*nx*ny elements: the same values in the per-destination
    slab layout of psb_propagate_ex (multi-GPU runs: each slab is one message of the frames -> kx-rows all-to-all).
    detector=(mask (nx, ny) float32 in shifted k order, out (L, Ptot, Ttot) float64, scratch (F*P, nx, ny) complex64):
    only sum_k |psi_k| * mask per (layer, probe, frame) is kept (HAADF without the exit-wave cube)."""
    F = t.shape[0]
    P = probes.shape[0]
    nx, ny, nz = plan.nx, plan.ny, plan.nz
    if work is None:
        work = torch.empty((F * P, nx, ny), dtype=torch.complex64, device=plan.device)
    L = _lib.lib()
    d = _lib.PropagateDesc()
    d.struct_bytes = C_sizeof(d)
    d.layer_every = layer_every
    d.n_frames, d.n_probes, d.nz, d.nx, d.ny = F, P, nz, nx, ny
    d.probes = probes.data_ptr()
    d.prop_x, d.prop_y = plan.prop_x.data_ptr(), plan.prop_y.data_ptr()
    d.psi_work = work.data_ptr()
    d.stream = _stream(plan.device)
    if t.dtype == torch.float32:                 # phase stack (build_transmission(phase=True))
        if t0 is None or t0.numel() < F * nx * ny:
            t0 = torch.empty((F, nx, ny), dtype=torch.complex64, device=plan.device)
        d.phase, d.t0_scratch = t.data_ptr(), t0.data_ptr()
    else:
        d.t = t.data_ptr()
    img = nx * ny
    if detector is not None:
        mask, out, scratch = detector
        assert mask.dtype == torch.float32 and mask.is_contiguous() and out.dtype == torch.float64 and out.is_contiguous()
        assert scratch.numel() >= F * P * img and out.shape[0] == layer_count(nz, layer_every)
        d.mode = 2
        d.det_mask, d.det_out, d.det_scratch = mask.data_ptr(), out.data_ptr(), scratch.data_ptr()
        d.det_stride_layer, d.det_stride_probe = out.stride(0), out.stride(1)
        d.frame0, d.probe0 = frame0, probe0
    elif wf_out is None:
        d.mode = 0
    elif slabs is not None:
        world, Lr, Tt, Pt = slabs
        assert wf_out.is_contiguous() and wf_out.numel() == Lr * Pt * Tt * img and Lr == layer_count(nz, layer_every)
        d.mode = 1
        d.wf_out = wf_out.data_ptr()
        d.slab_world, d.slab_layers, d.slab_frames, d.slab_probes = world, Lr, Tt, Pt
        d.frame0, d.probe0 = frame0, probe0
    else:
        Lr, Pt, Tt = wf_out.shape[:3]
        assert wf_out.is_contiguous() and Lr == layer_count(nz, layer_every)
        d.mode = 1
        d.wf_out = wf_out.data_ptr() + 8 * (probe0 * Tt * img + frame0 * img)
        d.stride_probe, d.stride_frame, d.stride_layer = Tt * img, img, Pt * Tt * img
    with _on(plan.device):
        _lib.check(L.psb_propagate_ex(d), "psb_propagate")
    if detector is not None:
        return detector[1]
    if wf_out is None:
        return work.view(F, P, nx, ny)
    return wf_out


def tacaw_intensity(wf_layer: torch.Tensor) -> torch.Tensor:
    """wf_layer (P, T, nx, ny) complex64 with contiguous (nx, ny) planes -> intensity (P, T, nx, ny) float32."""
    P, T, nx, ny = wf_layer.shape
    assert wf_layer.dtype == torch.complex64 and wf_layer.stride(3) == 1 and wf_layer.stride(2) == ny
    out = torch.empty((P, T, nx, ny), dtype=torch.float32, device=wf_layer.device)
    with _on(wf_layer.device):
        _lib.check(_lib.lib().psb_tacaw_intensity(wf_layer.data_ptr(), wf_layer.stride(0), wf_layer.stride(1), P, T,
                                                  nx * ny, out.data_ptr(), _stream(wf_layer.device)), "psb_tacaw_intensity")
    return out


def sum_pixels(x: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x (rows, npix) float32 or complex64 (|z| is summed), contiguous rows -> (rows,) float64."""
    rows, npix = x.shape
    assert x.stride(1) == 1
    out = torch.empty((rows,), dtype=torch.float64, device=x.device)
    m = mask.data_ptr() if mask is not None else None
    L = _lib.lib()
    fn = L.psb_sum_abs_pixels if x.dtype == torch.complex64 else L.psb_sum_pixels
    assert x.dtype in (torch.complex64, torch.float32)
    with _on(x.device):
        _lib.check(fn(x.data_ptr(), m, rows, x.stride(0), npix, out.data_ptr(), _stream(x.device)), "psb_sum_pixels")
    return out


def sum_frames(x: torch.Tensor) -> torch.Tensor:
    """x (G, T, npix) float32 contiguous -> (G, npix) float32 sums over T."""
    G, T, npix = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty((G, npix), dtype=torch.float32, device=x.device)
    with _on(x.device):
        _lib.check(_lib.lib().psb_sum_frames(x.data_ptr(), G, T, npix, out.data_ptr(), _stream(x.device)), "psb_sum_frames")
    return out


def _sm_count() -> int:
    return int(_lib.lib().psb_sm_count()) if not _lib.is_emulated() else 148


def _fill_images(max_imgs: int, tiles_per_img: int, slots: int) -> int:
    """Largest image count <= max_imgs whose tiles fill whole rounds of the persistent grids best: the
    slice-step kernels run `slots` tile pipelines side by side, so n images cost ceil(n*tiles/slots) rounds."""
    best, best_eff = max_imgs, 0.0
    for n in range(max_imgs, max(0, max_imgs // 2), -1):
        work = n * tiles_per_img
        eff = work / (-(-work // slots) * slots)
        if eff > best_eff + 1e-9:
            best, best_eff = n, eff
    return best


def _fewest_rounds(n_frames: int, fb_cap: int, tiles_per_frame: int, slots: int) -> int:
    """Frames per batch <= fb_cap that minimises the rounds of the persistent slice-step grids summed over all
    batches of the run (full batches plus the remainder); ties go to the larger batch.  500 frames of one 256 x 256
    probe on 148 SMs: 125 per batch = 4 x 7 rounds, where the balanced 5 x 100 costs 5 x 6."""
    best, best_rounds = fb_cap, None
    for fb in range(fb_cap, max(1, fb_cap // 2) - 1, -1):
        full, rem = divmod(n_frames, fb)
        rounds = full * -(-fb * tiles_per_frame // slots) + (-(-rem * tiles_per_frame // slots) if rem else 0)
        if best_rounds is None or rounds < best_rounds:
            best, best_rounds = fb, rounds
    return best


def batch_sizes(plan: SlicePlan, n_probes: int, n_frames: int):
    """(frames per batch, probes per sub-batch) so the psi batch stays L2-resident, the transmission buffer
    stays bounded, the image count fills the persistent slice-step grids, and batches are balanced."""
    img = plan.nx * plan.ny * 8
    per_frame_t = plan.nz * img
    max_imgs = max(1, (PSI_BATCH_BYTES_STREAMING if plan.nx * plan.ny >= 1024 * 1024 else PSI_BATCH_BYTES) // img)
    slots = (1 if plan.nx >= 1024 else 2) * _sm_count()          # CTAs of the column pass that run side by side
    tiles_per_img = max(1, plan.ny // (8 if plan.nx >= 512 else 16))
    if n_probes >= max_imgs:
        fb, pb = 1, _fill_images(max_imgs, tiles_per_img, slots)
        n_sub = -(-n_probes // pb)
        pb = -(-n_probes // n_sub)                    # balanced probe sub-batches
    else:
        fb_cap = max(1, max_imgs // n_probes)
        fb_cap = min(fb_cap, max(1, T_BATCH_BYTES // per_frame_t), max(1, 65535 // plan.nz), n_frames)
        if plan.device.type == "cuda":
            free, _ = torch.cuda.mem_get_info(plan.device)
            fb_cap = max(1, min(fb_cap, int(free * 0.5) // per_frame_t))
        fb = _fewest_rounds(n_frames, fb_cap, n_probes * tiles_per_img, slots)
        pb = n_probes
    return fb, min(pb, 65535 // max(1, fb))
