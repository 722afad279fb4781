#!/usr/bin/env python
"""Benchmark of the multislice + TACAW hot path (BASELINE.json metric: slice-steps/s).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the UNMODIFIED reference (torch CPU path) from baseline/_ref

Default workload = config C2 of BASELINE.json: plane-wave TACAW on 10 000-atom Si, 256 x 256 grid, 512 slices,
500 MD frames per GPU (weak scaling: frames are sharded by rank), full time-axis FFT to the THz spectrum.
`--workload` selects the other configurations (C1, C3, C4 = strong scaling over 2 000 frames, C5).
One "step" = potential build + propagation of every frame + exit FFT + [all-to-all] + TACAW time FFT.

The JSON line carries: value (inputs resident in HBM), e2e (public API with pinned-host inputs, H2D and
D2H inside the timed region), roofline of the slice-step kernels (+ the fp32-pipe floor of the same kernels),
cpu_baseline (the reference's torch path on the host cores, and the NumPy port beside it), clocks sampled during the
timed region, the number of kernels this library launched and -- on more than one GPU -- the outcome of the
N-GPU == 1-GPU parity check run before the warm-up.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # frames: per GPU for "weak" scaling, in total for "strong"; probes = n -> n x n probe_grid over the central half of the box
    "c2_si_256x256x512_500f_planewave": dict(sample="si", cells=(5, 5, 50), a=5.11, frames=500, grid=(256, 256, 512), seed=1,
                                             probes=0, aperture=0.0, layer_every=0, scaling="weak"),
    "c1_si_256x256x103_20f_planewave": dict(sample="si", cells=(5, 5, 10), a=5.11, frames=20, grid=(256, 256, 103), seed=0,
                                            probes=0, aperture=0.0, layer_every=0, scaling="weak"),
    "c3_hbn_512x512x67_100f_16x16": dict(sample="hbn", frames=100, grid=(512, 512, 67), seed=2, probes=16, aperture=30.0,
                                         layer_every=0, scaling="weak"),
    "c4_si_1024x1024x123_2000f_planewave": dict(sample="si", cells=(20, 20, 12), a=5.1175, frames=2000, grid=(1024, 1024, 123),
                                                seed=3, probes=0, aperture=0.0, layer_every=0, scaling="strong"),
    # C5 = 500 frames x 64 probes x 7 layers x 2 MB = 470 GB of exit waves: 63 frames (an eighth) per GPU
    "c5_hbn_512x512x67_63f_8x8_layers10": dict(sample="hbn", frames=63, grid=(512, 512, 67), seed=4, probes=8, aperture=30.0,
                                               layer_every=10, scaling="weak"),
}
ALIASES = {"c1": "c1_si_256x256x103_20f_planewave", "c2": "c2_si_256x256x512_500f_planewave",
           "c3": "c3_hbn_512x512x67_100f_16x16", "c4": "c4_si_1024x1024x123_2000f_planewave",
           "c5": "c5_hbn_512x512x67_63f_8x8_layers10"}
DEFAULT = "c2_si_256x256x512_500f_planewave"
VOLTAGE = 100e3

# fp32 issue floor of the fused slice step: SM-cycles per pixel and slice step that the packed-fp32
# instructions of the row + column kernels need when nothing else stalls, from their SASS (profiles/r2_sass_mix.txt) at the
# measured per-instruction cost (profiles/r2_ubench_fp32x2_operands.txt: the register file, not the fma pipe, sets it --
# FADD2 2.1 cycles per warp and scheduler, FFMA2 2.1 to 3.2 by distinct source registers).  Key: (line length, stack format).
FP32_SM_CYCLES_PER_PIXEL = {(256, "phase"): 0.499 + 0.478, (256, "c64"): 0.441 + 0.478, (512, "phase"): 0.617 + 0.597,
                            (512, "c64"): 0.560 + 0.597, (1024, "phase"): 0.622 + 0.601, (1024, "c64"): 0.564 + 0.601}


def make_traj(wl, n_frames, frame0=0):
    from pyslice_b200 import synthetic
    if wl["sample"] == "hbn":
        return synthetic.hbn_graphene_trajectory(n_frames=n_frames, seed=wl["seed"], frames=(frame0, frame0 + n_frames))
    return synthetic.silicon_trajectory(cells=wl["cells"], a=wl["a"], n_frames=n_frames, seed=wl["seed"],
                                        displacement="phonon", frames=(frame0, frame0 + n_frames))


def probe_positions(wl, box_matrix):
    if not wl["probes"]:
        return None
    from pyslice_b200.multislice.multislice import probe_grid
    lx, ly = box_matrix[0, 0], box_matrix[1, 1]
    n = wl["probes"]
    return [tuple(p) for p in probe_grid([0.25 * lx, 0.75 * lx], [0.25 * ly, 0.75 * ly], n, n)]


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled during the timed region (B200_PROFILING.md recipe: the
    nvidia-smi clocks line).  Source: NVML loaded into this process (the library nvidia-smi itself reads), one light
    query per 400 ms from a thread plus one when the timed region opens (each query can hold a driver lock that kernel
    launches of this process also take: 17 polls at 100 ms cost a 10-step run ~10 ms of idle GPU per step).  Any poller in ANOTHER process -- `nvidia-smi -lms` or an NVML child, at 100 to
    250 ms -- was measured to hold up this process's kernel launches at random, 1-17 ms of idle time per step in the
    device-timed arm (48 ms while nvidia-smi starts); the in-process thread does not.  Because runs with NVML loaded
    showed a noisier end-to-end arm, that arm is timed FIRST, before NVML is touched.
    Fallback when NVML cannot be loaded: an nvidia-smi child, started before the warm-up."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.first = [], None, index, 0
        self.nvml, self.handle, self.stop_flag, self.source = None, None, threading.Event(), None
        self.wake = threading.Event()

    def _nvml_open(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.nvml = pynvml

    def _nvml_loop(self):
        n = self.nvml
        masks = [("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap)]
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        while not self.stop_flag.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                bits = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                self.rows.append([sm, mx, 0.0] + ["Active" if bits & m else "Not Active" for _, m in masks])
            except Exception:
                pass
            self.wake.wait(0.4)          # every 400 ms, and at once when mark() opens the timed region
            self.wake.clear()

    def wait_ready(self, timeout=8.0):
        """block until the first sample arrived (start-up of the source is over)"""
        t0 = time.time()
        while (self.proc is not None or self.nvml is not None) and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.05)

    def mark(self):
        """the timed region starts here: earlier samples (warm-up) are not reported"""
        self.first = len(self.rows)
        self.wake.set()

    def start(self):
        try:
            self._nvml_open()
            self.source = "nvml"
            self.th = threading.Thread(target=self._nvml_loop, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None and self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        self.stop_flag.set()
        self.wake.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            self.th.join(timeout=2)
            try:
                self.nvml.nvmlShutdown()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if str(v).lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


# ---- CPU arms ------------------------------------------------------------------------------------------------------
def reference_root():
    """where the unmodified reference lives: baseline/_ref (tools/install_reference.py; travels to the GPU box), else the
    build container's checkout"""
    for cand in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("PYSLICE_REFERENCE", "/root/reference")):
        if os.path.exists(os.path.join(cand, "src", "multislice", "calculators.py")) and os.path.exists(os.path.join(cand, "kirkland.txt")):
            return cand
    return None


_PILOT = {}


def reference_sample_frames(wl, name, target_s=12.0):
    """frames of the bounded sample, sized on the box itself: a two-frame pilot run of the reference (it also warms
    torch's thread pool and the page cache) gives seconds per frame, and the sample is as many frames as fill ~target_s
    of CPU work (the GPU hosts of this pool run a 256^2 x 512 plane-wave frame in ~0.3 s on 16 threads, the build
    container takes 3.6 s on 8)"""
    key = (name, wl["frames"])
    if key not in _PILOT:
        reference_torch_step(wl, name, n_frames=2)                       # start-up costs (thread pool, imports, page cache)
        _, d = reference_torch_step(wl, name, n_frames=min(6, max(2, wl["frames"])))
        _PILOT[key] = d["seconds"] / max(2, min(6, wl["frames"]))
    per_frame = _PILOT[key]
    return int(max(2, min(96, wl["frames"], round(target_s / max(per_frame, 1e-3)))))


def reference_torch_step(wl, name, n_frames=None, n_probes_cap=16):
    """One bounded sample of the workload through the REFERENCE ITSELF: MultisliceCalculator(force_cpu=True).setup().run()
    (reference src/multislice/calculators.py:41,96,163) + TACAWData (src/postprocessing/tacaw_data.py:38), full slice
    stack, all host threads torch can use, fresh temp cwd (its frame cache key ignores the positions).
    Returns (slice-steps/s, description dict) or raises when the reference is not installed."""
    import torch
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not installed: run tools/install_reference.py in the build container")
    if root not in sys.path:
        sys.path.insert(0, root)
    from src.multislice.calculators import MultisliceCalculator as RefCalc        # noqa: E402  (the reference's modules)
    from src.multislice.trajectory import Trajectory as RefTraj                   # noqa: E402
    from src.postprocessing.tacaw_data import TACAWData as RefTACAW               # noqa: E402
    import logging
    logging.getLogger("src.multislice.calculators").setLevel(logging.WARNING)
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(cores)
    n = n_frames or reference_sample_frames(wl, name)
    traj = make_traj(wl, max(n, 2))            # TACAWData needs two time points for its frequency axis
    n = traj.n_frames
    pp = probe_positions(wl, traj.box_matrix)
    if pp is not None and len(pp) > n_probes_cap:
        pp = pp[:: len(pp) // n_probes_cap][:n_probes_cap]
    P = len(pp) if pp is not None else 1
    ref_traj = RefTraj(atom_types=traj.atom_types, positions=traj.positions, velocities=traj.velocities,
                       box_matrix=traj.box_matrix, timestep=traj.timestep)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="pyslice_ref_")
    os.chdir(tmp)
    try:
        t0 = time.perf_counter()
        calc = RefCalc(force_cpu=True)
        calc.setup(ref_traj, aperture=wl["aperture"], voltage_eV=VOLTAGE, probe_positions=pp)      # cleanup_temp_files=True hits a NameError in the reference (calculators.py:237)
        wf = calc.run()
        t1 = time.perf_counter()
        tac = RefTACAW(wf)
        _ = tac.spectrum()
        t2 = time.perf_counter()
    finally:
        os.chdir(cwd)
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    nz = calc.nz
    assert (calc.nx, calc.ny, nz) == tuple(wl["grid"]), (calc.nx, calc.ny, nz)
    steps = P * n * nz
    dt = t2 - t0
    return steps / dt, {
        "value": steps / dt, "unit": "slice-steps/s", "cores": cores, "kind": "reference",
        "impl": "reference-torch: unmodified h-walk/PySlice MultisliceCalculator(force_cpu=True).setup().run() + TACAWData, "
                "torch CPU complex128, imported from " + os.path.relpath(root, ROOT),
        "sample": f"{n} of {wl['frames']} frames, {P} probe(s), all {nz} slices of workload {name}: run() {t1 - t0:.1f} s "
                  f"+ TACAWData/spectrum {t2 - t1:.2f} s, torch.set_num_threads({cores})",
        "slice_steps_in_sample": steps, "seconds": dt, "torch_threads": cores}


def port_step(wl, z_fraction=1.0):
    """The oracle (NumPy/SciPy port of the reference's torch path, float64) on the host cores: reported beside the
    reference's own figure as a second, labelled baseline (it is the faster of the two)."""
    from oracle import pyslice_oracle as orc
    cores = os.cpu_count() or 1
    n = max(2, min(cores, 16))
    cz = max(2, int(round(wl["cells"][2] * z_fraction))) if wl["sample"] == "si" else None
    thin = dict(wl, cells=(wl["cells"][0], wl["cells"][1], cz)) if cz else wl
    traj = make_traj(thin, n)
    pp = probe_positions(wl, traj.box_matrix)
    if pp is not None:
        pp = pp[:: max(1, len(pp) // 4)][:4]
    t0 = time.time()
    wf, grid = orc.multislice_run(traj.positions, traj.atom_types, traj.box_matrix, aperture=wl["aperture"], voltage_eV=VOLTAGE,
                                  probe_positions=pp, frame_threads=n, workers=1)
    orc.tacaw_intensity(wf[..., 0], np.arange(n) * traj.timestep)
    dt = time.time() - t0
    nz = len(grid["zs"])
    P = wf.shape[0]
    return {"value": P * n * nz / dt, "unit": "slice-steps/s", "cores": n, "kind": "port",
            "sample": f"{n} frames x {P} probe(s) x {nz} slices, {n} frame threads, NumPy/SciPy float64, {dt:.1f} s"}


def cpu_baseline(wl, name):
    """cpu_baseline object of the GPU arm's line (rank 0, N = 1): the reference's torch path on a bounded sample; the
    NumPy port as a second figure under "port"."""
    try:
        _, base = reference_torch_step(wl, name)
    except Exception as e:                                     # reference tree absent: the port is all there is
        base = port_step(wl)
        base["note"] = f"reference not runnable here ({type(e).__name__}: {e}); NumPy port instead"
        return base
    try:
        base["port"] = port_step(wl, z_fraction=0.25)
    except Exception as e:                                     # pragma: no cover
        base["port"] = {"error": str(e)}
    return base


L2_NOTE = ("GPU arm: inputs larger than L2, no explicit flush -- every timed step streams the positions and a multi-GB "
           "transmission stack per frame batch through HBM; only the psi batch (<= 80 MB) is L2-resident by design "
           "(1024-point grids: psi streams from HBM too, batches of up to 320 MB)")


def config_block(wl, name, world, counts, A, P):
    """the `config` object, identical in the GPU arm and the reference arm of one invocation"""
    nx, ny, nz = wl["grid"]
    return {"workload": name, "grid": [nx, ny, nz], "atoms": A, "frames_per_gpu": max(counts), "frames_total": sum(counts),
            "probes": P, "aperture_mrad": wl["aperture"], "layer_every": wl["layer_every"], "voltage_eV": VOLTAGE,
            "parallelism": f"frames sharded over {world} GPU(s), all-to-all to kx rows", "l2": L2_NOTE}


def frame_counts(wl, world):
    if wl["scaling"] == "strong":
        return [wl["frames"] // world + (1 if r < wl["frames"] % world else 0) for r in range(world)]
    return [wl["frames"]] * world


def run_reference(args, wl, name):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, base = [], None
    total = args.warmup + args.steps
    # every step is one bounded sample (~12 s of the reference's CPU work); beyond 8 steps the sample shrinks so that the
    # whole run stays within a few minutes
    try:
        n = reference_sample_frames(wl, name, target_s=12.0 if total <= 8 else max(2.0, 100.0 / total))
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"}))
        return
    try:
        for i in range(total):
            v, base = reference_torch_step(wl, name, n_frames=n)
            if i >= args.warmup:
                vals.append(v)
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"}))
        return
    v = float(np.median(vals))
    base["value"] = v
    base["all_steps"] = vals
    P = wl["probes"] ** 2 if wl["probes"] else 1
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = config_block(wl, name, world, frame_counts(wl, world), int(make_traj(wl, 1).n_atoms), P)
    base["note"] = "reference arm: one bounded sample of the workload in `config` per step (see sample), median of the steps"
    line = {"impl": "reference", "metric": "slice-steps/sec (probe*frame*slice)", "value": v, "unit": "slice-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * base["slice_steps_in_sample"] / v, "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "c128 (torch CPU)", "data": "synthetic", "config": cfg, "cpu_baseline": base,
            "e2e": {"value": v, "unit": "slice-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---- N-GPU == 1-GPU ------------------------------------------------------------------------------------------------
def multi_gpu_parity(dev, rank, world):
    """A small frame-sharded job over the real process group (NCCL on the GPU box) against the same job computed
    unsharded by every rank: intensity rows bit for bit, reducers to float64 round-off (reference: the sequential frame
    loop of src/multislice/calculators.py:172 + src/postprocessing/tacaw_data.py:89-104 -- the single-GPU result IS the
    oracle of the sharded one).  256 x 256 x 9 slices, 8*world + 3 frames (ragged).  Returns "bitwise" or raises."""
    import torch
    import torch.distributed as dist
    from pyslice_b200 import synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.postprocessing.tacaw_data import TACAWData
    T = 8 * world + 3
    traj = synthetic.random_trajectory(n_atoms=400, box=(25.55, 25.55, 4.1), n_frames=T, seed=41, types=(6, 14))
    calc = MultisliceCalculator(device=dev)
    calc.setup(traj, aperture=0.0, voltage_eV=VOLTAGE)                          # sharded: world > 1
    assert calc.shard is not None and sum(calc.shard.counts) == T
    wf = calc.run()
    tac = TACAWData(wf)
    single = MultisliceCalculator(device=dev)
    single.setup(traj, aperture=0.0, voltage_eV=VOLTAGE, shard_frames=False)
    tac1 = TACAWData(single.run())
    r0, r1 = tac.row_range
    ok = bool(torch.equal(tac.intensity, tac1.intensity[:, :, r0:r1]))
    s, s1 = tac.spectrum(), tac1.spectrum()                                     # collective inside: every rank calls
    d, d1 = tac.diffraction(), tac1.diffraction()
    ok = ok and bool(np.allclose(s, s1, rtol=1e-12, atol=0)) and bool(np.allclose(d, d1, rtol=1e-6, atol=0))
    ok = ok and bool(np.abs(s1).max() > 0)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) != 1:
        raise AssertionError(f"rank {rank}: sharded result differs from the single-GPU result")
    return "bitwise"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT, choices=list(WORKLOADS) + list(ALIASES))
    ap.add_argument("--frames", type=int, default=0, help="override the workload's frame count (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end arm (profiling runs)")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample clocks (diagnostic: NVML polls can hold up launches)")
    args = ap.parse_args()
    name = ALIASES.get(args.workload, args.workload)
    wl = dict(WORKLOADS[name])
    if args.frames:
        wl["frames"] = args.frames
    if args.impl == "reference":
        return run_reference(args, wl, name)

    import torch
    import torch.distributed as dist
    from pyslice_b200 import engine
    from pyslice_b200.multislice.calculators import FrameShard, MultisliceCalculator
    from pyslice_b200.multislice.trajectory import Trajectory
    from pyslice_b200.postprocessing.tacaw_data import TACAWData

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    parity = multi_gpu_parity(dev, rank, world) if world > 1 else None

    nx, ny, nz = wl["grid"]
    counts = frame_counts(wl, world)
    F = counts[rank]
    frame0 = sum(counts[:rank])
    T_total = sum(counts)
    # every rank generates only its own block of the same global trajectory
    local_traj = make_traj(wl, F, frame0=frame0)
    A = local_traj.n_atoms
    pp = probe_positions(wl, local_traj.box_matrix)
    P = len(pp) if pp is not None else 1

    class ShardedCalc(MultisliceCalculator):
        """the public calculator, fed this rank's frame block directly (the global trajectory is never
        materialised on one host): n_frames / shard describe the global run"""
        def _local_frames(self):
            return 0, F

    def setup_calc(traj):
        calc = ShardedCalc(device=dev)
        calc.setup(traj, aperture=wl["aperture"], voltage_eV=VOLTAGE, probe_positions=pp, layer_every=wl["layer_every"],
                   shard_frames=False)
        if world > 1:
            calc.shard = FrameShard(rank, world, counts)
            calc.n_frames = T_total
        return calc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rows = nx // world + (1 if rank < nx % world else 0)

    def result_to_host(tac, out_host):
        """the step's result leaves the device: the intensity cube of a plane-wave TACAW run (C1, C2, C4), the per-probe
        spectra and the mean diffraction pattern of a STEM run (the cube is tens of GB there)"""
        if P == 1:
            out_host.copy_(tac.intensity, non_blocking=True)
            torch.cuda.synchronize()
            return out_host.numel() * 4
        s = tac._sum_k()
        d = tac.diffraction()
        return s.nbytes + d.nbytes

    # ---------------- arm 2 (timed first, see ClockSampler): end to end through the public API with host buffers ----
    ms_e2e, d2h = None, 0
    if not args.no_e2e:
        pinned = torch.empty(local_traj.positions.shape, dtype=torch.float64).pin_memory()
        pinned.numpy()[...] = local_traj.positions
        host_traj = Trajectory(local_traj.atom_types, pinned.numpy(), np.zeros((F, A, 3)),
                               local_traj.box_matrix, local_traj.timestep)
        out_host = torch.empty((1, T_total, rows, ny), dtype=torch.float32).pin_memory() if P == 1 else None

        def e2e_step():
            c = setup_calc(host_traj)
            wf = c.run()
            tac = TACAWData(wf)
            return result_to_host(tac, out_host)

        for _ in range(args.warmup):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            d2h = e2e_step()
        barrier()
        ms_e2e = (time.perf_counter() - t0) * 1e3
        del pinned, host_traj, out_host
        torch.cuda.empty_cache()

    # ---------------- arm 1: inputs resident in HBM ------------------------------------------
    pos_dev = torch.from_numpy(local_traj.positions).to(dev)
    dev_traj = Trajectory.__new__(Trajectory)
    dev_traj.atom_types, dev_traj.positions, dev_traj.velocities = local_traj.atom_types, pos_dev, None
    dev_traj.box_matrix, dev_traj.timestep = local_traj.box_matrix, local_traj.timestep
    calc = setup_calc(dev_traj)
    assert (calc.nx, calc.ny, calc.nz) == (nx, ny, nz), (calc.nx, calc.ny, calc.nz)
    timer = engine.PhaseTimer(dev)

    def device_step(tm=None):
        wf = calc.run(timer=tm)
        wf._timer = tm
        tac = TACAWData(wf)
        return tac

    # the clock sampler is started BEFORE the warm-up (its start-up must not land inside the timed region); only the
    # samples taken during the timed region are reported
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        sampler.start()
        sampler.wait_ready()
    for _ in range(args.warmup):
        tac = device_step()
        del tac
    barrier()
    sampler.mark()
    l0 = engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tac = None
    marks = []
    for _ in range(args.steps):
        tac = device_step(timer)        # the previous result stays alive until the new one exists (as a caller's loop would)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append(ev)
    e1.record()
    barrier()
    launches = engine.launch_count() - l0
    ms_dev = e0.elapsed_time(e1)
    step_ms = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    phases = timer.totals()
    clocks = sampler.stop() if rank == 0 and not args.no_clocks else None
    spectrum = tac.spectrum()          # touches the result (and checks the reducers run)
    assert np.isfinite(spectrum).all() and np.abs(spectrum).max() > 0
    del tac

    keys = ["propagate", "potential", "all_to_all", "tacaw"]
    vals = [ms_dev, ms_e2e or 0.0] + [phases.get(k, 0.0) for k in keys]
    if world > 1:
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = t.tolist()
    ms_dev, ms_e2e_max = vals[0], vals[1]
    prop_ms, pot_ms, a2a_ms, tacaw_ms = vals[2:6]

    if rank == 0:
        slice_steps = P * T_total * nz                       # probes * frames * slices, whole job
        value = slice_steps * args.steps / (ms_dev * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        b_ss = nx * ny * (16 + 4 / P)                        # psi r+w (c64) + fp32 phase shared by P probes
        per_gpu_steps = P * max(counts) * nz * args.steps
        achieved = per_gpu_steps * b_ss / (prop_ms * 1e-3) / 1e9 if prop_ms > 0 else None
        traffic_doc = {}
        try:
            traffic_doc = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        except Exception:
            pass
        traffic = traffic_doc.get("by_grid", {}).get(f"{nx}x{ny}", {}).get("dram_bytes_per_launch_pair",
                                                                          traffic_doc.get("slice_step_dram_bytes_per_launch") if nx == 256 else None)
        fb, pb = engine.batch_sizes(calc._plan, P, max(F, 1))
        by_grid = traffic_doc.get("by_grid", {}).get(f"{nx}x{ny}")
        if traffic is not None and by_grid and by_grid.get("images_per_launch"):
            traffic = traffic * (fb * pb) / by_grid["images_per_launch"]          # captured per launch pair of N images; scale to this run's batch
        cfg = config_block(wl, name, world, counts, A, P)
        cyc = FP32_SM_CYCLES_PER_PIXEL.get((max(nx, ny), "phase" if P == 1 else "c64")) if nx == ny else None
        sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_floor = None
        if cyc:
            floor_steps = 148 * sm_clock * 1e6 / (cyc * nx * ny)        # slice-steps/s if the fp32 instructions alone filled the SMs
            fp32_floor = {"sm_cycles_per_pixel": cyc, "slice_steps_per_s": floor_steps,
                          "frac_of_floor": (per_gpu_steps / (prop_ms * 1e-3)) / floor_steps if prop_ms > 0 else None,
                          "hbm_roofline_slice_steps_per_s": peak * 1e9 / b_ss,
                          "max_frac_of_hbm_roofline": floor_steps / (peak * 1e9 / b_ss),
                          "note": "register-file-limited issue time of the kernels' own packed-fp32 instructions (DESIGN.md 4.1, "
                                  "profiles/r2_sass_mix.txt); roofline.frac cannot exceed max_frac_of_hbm_roofline with these transforms"}
        line = {
            "metric": "slice-steps/sec (probe*frame*slice)", "value": value, "unit": "slice-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "c64 (fp32 complex)",
            "data": "synthetic", "config": cfg,
            "run": {"frames_per_batch": fb, "tacaw_wall_ms": ms_dev / args.steps, "step_ms": [round(x, 2) for x in step_ms]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "phases_ms_per_step": {"potential": pot_ms / args.steps, "propagate_incl_exit_fft": prop_ms / args.steps,
                                   "all_to_all": a2a_ms / args.steps, "tacaw": tacaw_ms / args.steps,
                                   "other": (ms_dev - pot_ms - prop_ms - a2a_ms - tacaw_ms) / args.steps},
            "roofline": {"bound": "hbm", "kernel": f"slice-step = fused row pass + column pass at {nx} x {ny} "
                                                   f"({'float32 phase stack' if P == 1 else 'complex64 transmission stack'})",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if achieved else None, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "algorithmic_bytes_per_slice_step": b_ss, "fp32_floor": fp32_floor,
                         "traffic_note": "ncu dram bytes per launch pair (row + column pass), see profiles/roofline_traffic.json"},
        }
        if ms_e2e is not None:
            line["e2e"] = {"value": slice_steps * args.steps / (ms_e2e_max * 1e-3), "unit": "slice-steps/s",
                           "ms_per_step": ms_e2e_max / args.steps,
                           "h2d_bytes_per_step": int(F * A * 3 * 8), "d2h_bytes_per_step": int(d2h),
                           "api": "MultisliceCalculator.setup()/run() + TACAWData(wf), pinned host positions in, " +
                                  ("intensity cube out" if P == 1 else "per-probe spectra + diffraction pattern out")}
        if a2a_ms > 0 and world > 1:
            sent = 8.0 * P * F * nx * ny * (world - 1) / world * args.steps         # bytes this rank sends per run
            line["all_to_all"] = {"ms_per_step": a2a_ms / args.steps, "bytes_sent_per_rank_per_step": sent / args.steps,
                                  "gb_per_s_per_rank": sent / (a2a_ms * 1e-3) / 1e9, "backend": "nccl all_to_all_single"}
        if parity is not None:
            line["multi_gpu_parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl, name)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
