"""Time psb_bin_atoms + psb_build_transmission on the C2 geometry for several scratch (chunk) sizes.
usage: python tools/microbench_potential.py [frames] [scratch_MB ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyslice_b200 import engine, hostmath, synthetic, _lib
if os.environ.get("PSB_VARIANT_LIB"):          # tuning experiments: time an alternative build of libpsb
    _lib._lib = _lib.load(os.path.abspath(os.environ["PSB_VARIANT_LIB"]))
    print("variant library:", os.environ["PSB_VARIANT_LIB"])

F = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sizes = [int(x) for x in sys.argv[2:]] or [32, 64, 96]
if os.environ.get("PSB_GEOM") == "c4":          # 38 400 atoms, 1024 x 1024 x 123
    traj = synthetic.silicon_trajectory(cells=(20, 20, 12), a=5.1175, n_frames=F, seed=3, displacement="phonon")
else:
    traj = synthetic.silicon_trajectory(cells=(5, 5, int(os.environ.get("PSB_ZCELLS", "50"))), a=5.11, n_frames=F, seed=1, displacement="phonon")
xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
pos = torch.from_numpy(traj.positions).cuda()
PHASE = os.environ.get("PSB_PHASE") == "1"      # stack as float32 phases (single-probe format)
engine.set_sf_mode(int(os.environ.get("PSB_SF_MODE", "0")))      # 0 auto, 1 direct sum, 2 NUFFT along x
print("sf mode", os.environ.get("PSB_SF_MODE", "0"))
tbuf = torch.empty((F, plan.nz, plan.nx, plan.ny), dtype=torch.complex64, device="cuda")      # large enough for either format
tphase = tbuf.view(torch.float32).reshape(-1)[:tbuf.numel()].view(F, plan.nz, plan.nx, plan.ny)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for fast in [int(x) for x in os.environ.get('PSB_LEVELS', '2,1,0').split(',')]:
    engine.set_fast_path(fast)
    for mb in sizes:
        engine.SCRATCH_BYTES = mb << 20
        if os.environ.get('PSB_CHUNK_PAIRS') == '1':          # sizes are slice-pair images per chunk instead of MB
            engine.SCRATCH_BYTES = mb * 8 * plan.nx * plan.ny
        for _ in range(2):
            engine.build_transmission(plan, pos, out=tphase if PHASE and fast > 0 else tbuf, phase=PHASE and fast > 0)
        torch.cuda.synchronize()
        a.record()
        for _ in range(3):
            engine.build_transmission(plan, pos, out=tphase if PHASE and fast > 0 else tbuf, phase=PHASE and fast > 0)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"{'level %d' % fast} scratch {mb:4d} MB  F={F}: {ms:8.3f} ms  {1e3*ms/(F*plan.nz):6.3f} us per slice  "
              f"({F*plan.nz/ms/1e3:6.3f} M slices/s)", flush=True)
    # binning alone
    for _ in range(2):
        engine.bin_atoms(plan, pos)
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        engine.bin_atoms(plan, pos)
    b.record(); torch.cuda.synchronize()
    print(f"   bin_atoms alone: {a.elapsed_time(b)/5:7.3f} ms", flush=True)
engine.set_fast_path(True)
