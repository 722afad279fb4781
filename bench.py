#!/usr/bin/env python
"""Benchmark of the multislice + TACAW hot path (BASELINE.json metric: slice-steps/s).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the CPU port of the reference path (oracle), all host threads

Workload (config C2 of BASELINE.json): plane-wave TACAW on 10 000-atom Si, 256 x 256 grid, 512 slices,
500 MD frames per GPU (weak scaling: frames are sharded by rank), full time-axis FFT to the THz spectrum.
One "step" = potential build + propagation of every frame + exit FFT + [all-to-all] + TACAW time FFT.

The JSON line carries: value (inputs resident in HBM), e2e (public API with pinned-host inputs, H2D and
D2H inside the timed region), roofline of the slice-step kernels, cpu_baseline (oracle on host cores),
clocks sampled during the timed region and the number of kernels this library launched.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (cells, lattice a, frames per GPU, grid, slices)
    "c2_si_256x256x512_500f_planewave": dict(cells=(5, 5, 50), a=5.11, frames=500, grid=(256, 256, 512), seed=1),
    "c1_si_256x256x103_20f_planewave": dict(cells=(5, 5, 10), a=5.11, frames=20, grid=(256, 256, 103), seed=0),
}
DEFAULT = "c2_si_256x256x512_500f_planewave"
VOLTAGE = 100e3


def make_traj(wl, n_frames, frame0=0):
    from pyslice_b200 import synthetic
    return synthetic.silicon_trajectory(cells=wl["cells"], a=wl["a"], n_frames=n_frames, seed=wl["seed"],
                                        displacement="phonon", frames=(frame0, frame0 + n_frames))


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


def cpu_baseline(wl, z_fraction=1.0):
    """The oracle (NumPy/SciPy port of the reference's torch path, float64) on the host cores: a bounded
    sample of the same workload -- `cores` frames in parallel threads, full slice stack each (z_fraction < 1: a
    thinner sample of the same crystal, for runs of many reference steps; the rate per slice-step is unchanged)."""
    from oracle import pyslice_oracle as orc
    cores = os.cpu_count() or 1
    n = max(1, min(cores, 16))
    cz = max(2, int(round(wl["cells"][2] * z_fraction)))
    thin = dict(wl, cells=(wl["cells"][0], wl["cells"][1], cz))
    traj = make_traj(thin, n)
    t0 = time.time()
    wf, grid = orc.multislice_run(traj.positions, traj.atom_types, traj.box_matrix, aperture=0.0, voltage_eV=VOLTAGE,
                                  frame_threads=n, workers=1)
    inten, _ = orc.tacaw_intensity(wf[..., 0], np.arange(n) * traj.timestep) if n > 1 else (None, None)
    dt = time.time() - t0
    nz = len(grid["zs"])
    return {"value": n * nz / dt, "unit": "slice-steps/s", "cores": n, "kind": "port", "slices": nz,
            "sample": f"{n} of {wl['frames']} frames ({nz} of {wl['grid'][2]} slices, potential + propagation + exit FFT + TACAW), "
                      f"{n} frame threads, float64, {dt:.1f} s"}


def run_reference(args, wl, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, base = [], None
    # every step is a bounded sample (~13 s at full thickness on 16 threads); beyond 8 steps in all the samples get
    # thinner so that the whole run stays within a few minutes
    total = args.warmup + args.steps
    zf = 1.0 if total <= 8 else 8.0 / total
    for i in range(total):
        base = cpu_baseline(wl, zf)
        if i >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    nz = wl["grid"][2]
    line = {"impl": "reference", "metric": "slice-steps/sec (probe*frame*slice)", "value": v, "unit": "slice-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * base["cores"] * base["slices"] / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "note": "CPU port of the reference path (oracle/), bounded frame sample per step"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "slice-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT, choices=list(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="override frames per GPU (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.frames:
        wl["frames"] = args.frames
    if args.impl == "reference":
        return run_reference(args, wl, args.workload)

    import torch
    import torch.distributed as dist
    from pyslice_b200 import engine
    from pyslice_b200.multislice.calculators import MultisliceCalculator
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

    F = wl["frames"]
    nx, ny, nz = wl["grid"]
    T_total = F * world
    # every rank generates only its own block of the same global trajectory (weak scaling in frames)
    local_traj = make_traj(wl, F, frame0=rank * F)
    A = local_traj.n_atoms

    class ShardedCalc(MultisliceCalculator):
        """the public calculator, fed this rank's frame block directly (the global trajectory is never
        materialised on one host): n_frames / shard describe the global run"""
        def _local_frames(self):
            return 0, F

    def setup_calc(traj):
        calc = ShardedCalc(device=dev)
        calc.setup(traj, aperture=0.0, voltage_eV=VOLTAGE, shard_frames=False)
        if world > 1:
            from pyslice_b200.multislice.calculators import FrameShard
            calc.shard = FrameShard(rank, world, [F] * world)
            calc.n_frames = T_total
        return calc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- arm 2 (timed first, see ClockSampler): end to end through the public API with host buffers ----
    pinned = torch.empty(local_traj.positions.shape, dtype=torch.float64).pin_memory()
    pinned.numpy()[...] = local_traj.positions
    host_traj = Trajectory(local_traj.atom_types, pinned.numpy(), np.zeros((F, A, 3)), local_traj.box_matrix,
                           local_traj.timestep)
    rows = nx // world + (1 if rank < nx % world else 0)
    out_host = torch.empty((1, T_total, rows, ny), dtype=torch.float32).pin_memory()

    def e2e_step():
        c = setup_calc(host_traj)
        wf = c.run()
        tac = TACAWData(wf)
        out_host.copy_(tac.intensity, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3

    # ---------------- arm 1: inputs resident in HBM ------------------------------------------
    pos_dev = torch.from_numpy(local_traj.positions).to(dev)
    dev_traj = Trajectory.__new__(Trajectory)
    dev_traj.atom_types, dev_traj.positions, dev_traj.velocities = local_traj.atom_types, pos_dev, None
    dev_traj.box_matrix, dev_traj.timestep = local_traj.box_matrix, local_traj.timestep
    calc = setup_calc(dev_traj)
    assert (calc.nx, calc.ny, calc.nz) == (nx, ny, nz)
    timer = engine.PhaseTimer(dev)

    def device_step(tm=None):
        wf = calc.run(timer=tm)
        tac = TACAWData(wf)
        return tac

    # the clock sampler is started BEFORE the warm-up (its start-up must not land inside the timed region); only the
    # samples taken during the timed region are reported
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler.mark()
    l0 = engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tac = device_step(timer)
    e1.record()
    barrier()
    launches = engine.launch_count() - l0
    ms_dev = e0.elapsed_time(e1)
    phases = timer.totals()
    clocks = sampler.stop() if rank == 0 else None
    spectrum = tac.spectrum()          # touches the result (and checks the reducers run)
    assert np.isfinite(spectrum).all()
    del tac

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e, phases.get("propagate", 0.0), phases.get("potential", 0.0)],
                         dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, prop_ms, pot_ms = t.tolist()
    else:
        prop_ms, pot_ms = phases.get("propagate", 0.0), phases.get("potential", 0.0)

    if rank == 0:
        slice_steps = 1 * T_total * nz                       # probes * frames * slices, whole job
        value = slice_steps * args.steps / (ms_dev * 1e-3)
        e2e_value = slice_steps * args.steps / (ms_e2e * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        b_ss = nx * ny * (16 + 4 / 1)                        # psi r+w (c64) + fp32 phase, B = 1 probe
        per_gpu_steps = F * nz * args.steps
        achieved = per_gpu_steps * b_ss / (prop_ms * 1e-3) / 1e9 if prop_ms > 0 else None
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("slice_step_dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": "slice-steps/sec (probe*frame*slice)", "value": value, "unit": "slice-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c64 (fp32 complex)",
            "data": "synthetic",
            "config": {"workload": args.workload, "grid": [nx, ny, nz], "atoms": A, "frames_per_gpu": F,
                       "frames_total": T_total, "probes": 1, "voltage_eV": VOLTAGE,
                       "l2": "inputs larger than L2: every timed step streams the positions and a >= 30 GB transmission stack per "
                             "frame batch through HBM; only the psi batch (<= 80 MB) is L2-resident by design; no explicit flush",
                       "frames_per_batch": engine.batch_sizes(calc._plan, 1, F)[0],
                       "tacaw_wall_ms": ms_dev / args.steps, "parallelism": f"frames sharded over {world} GPU(s), all-to-all to kx rows"},
            "e2e": {"value": e2e_value, "unit": "slice-steps/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(F * A * 3 * 8), "d2h_bytes_per_step": int(T_total * rows * ny * 4),
                    "api": "MultisliceCalculator.setup()/run() + TACAWData(wf), pinned host positions in, intensity out"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "phases_ms_per_step": {"potential": pot_ms / args.steps, "propagate_incl_exit_fft": prop_ms / args.steps,
                                   "other_incl_tacaw": (ms_dev - pot_ms - prop_ms) / args.steps},
            "roofline": {"bound": "hbm", "kernel": "slice-step = fast_rows_kernel<256,3> + fast_cols_kernel<256,256,0> (psb_propagate_phase)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if achieved else None, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "algorithmic_bytes_per_slice_step": b_ss,
                         "traffic_note": "ncu dram bytes per launch pair (row + column pass, cold L2 under ncu), see profiles/roofline_traffic.json"},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
