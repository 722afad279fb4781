"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (h-walk/PySlice torch path, CPU, complex128).

Build-container only: needs the reference checkout (default /root/reference, override with
PYSLICE_REFERENCE).  The reference caches frames under ./psi_data with a key that ignores atom
positions (reference src/multislice/calculators.py:78-94), so every run happens in a fresh temp cwd.

    python tests/golden/make_golden.py

Fixtures (inputs are regenerated from seeds by the tests via pyslice_b200.synthetic):
  small64.npz    64x64x9 grid, 3 element types, atoms on/near slice bounds, 3 frames
                 -> Potential.array (frame 0), plane-wave run(), 2x2-probe 30 mrad run(), HAADF
  tacaw48.npz    48x48 grid (non power of two), 12 frames -> run(), TACAWData intensity + reducers
  probe_kat.npz  the reference's own 00_probe.py recipe (501x491 grid, 1/3/5/15/30 mrad), subsampled
  si_c1.npz      config C1 geometry (256x256x103, 2000 Si atoms), frame 0: potential checksum planes,
                 exit wave (plane wave) -- stored as complex64 to stay small
"""
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("PYSLICE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, REF)

import torch  # noqa: E402

from src.multislice.calculators import MultisliceCalculator  # noqa: E402
from src.multislice.multislice import Probe, probe_grid  # noqa: E402
from src.multislice.potentials import Potential, gridFromTrajectory  # noqa: E402
from src.multislice.trajectory import Trajectory as RefTrajectory  # noqa: E402
from src.postprocessing.haadf_data import HAADFData  # noqa: E402
from src.postprocessing.tacaw_data import TACAWData  # noqa: E402

from pyslice_b200 import synthetic  # noqa: E402


def fresh_cwd():
    os.chdir(tempfile.mkdtemp(prefix="pyslice_ref_"))


def to_ref(traj):
    return RefTrajectory(atom_types=traj.atom_types, positions=traj.positions,
                         velocities=traj.velocities, box_matrix=traj.box_matrix, timestep=traj.timestep)


def ref_run(traj, **kw):
    fresh_cwd()
    calc = MultisliceCalculator(force_cpu=True)
    calc.setup(to_ref(traj), **kw)
    wf = calc.run()
    return calc, wf


def names_for(calc, atom_types):
    return [calc.element_map.get(int(z), int(z)) for z in atom_types]


def small64():
    traj = synthetic.random_trajectory(n_atoms=200, box=(6.35, 6.35, 4.1), n_frames=3, seed=11, stray=True)
    calc, wf_pw = ref_run(traj, aperture=0.0, voltage_eV=100e3)
    pot = Potential(calc.xs, calc.ys, calc.zs, traj.positions[0], names_for(calc, traj.atom_types),
                    kind="kirkland", device="cpu")
    xy = probe_grid([1.0, 4.0], [1.5, 5.0], 2, 2)
    calc2, wf_cb = ref_run(traj, aperture=30.0, voltage_eV=100e3, probe_positions=xy)
    adf = HAADFData(wf_cb).calculateADF(collection_angle=45)
    np.savez_compressed(
        os.path.join(HERE, "small64.npz"),
        potential0=pot.array.numpy(), wf_plane=wf_pw.wavefunction_data.numpy(),
        wf_probes=wf_cb.wavefunction_data.numpy(), probe_xy=xy,
        base_probe=calc2.base_probe.array.numpy(), adf=np.asarray(adf),
        kxs=wf_pw.kxs.numpy(), kys=wf_pw.kys.numpy(), time=wf_pw.time,
        xs=calc.xs, ys=calc.ys, zs=calc.zs)


def tacaw48():
    traj = synthetic.random_trajectory(n_atoms=120, box=(4.75, 4.75, 3.2), n_frames=12, seed=5, types=(5, 7))
    calc, wf = ref_run(traj, aperture=0.0, voltage_eV=100e3)
    tac = TACAWData(wf)
    inten = tac.intensity.numpy()
    kx_path = np.linspace(-2, 2, 7)
    ky_path = np.linspace(0, 3, 7)
    mask = (np.abs(np.asarray(wf.kxs))[:, None] < 2.0) & (np.abs(np.asarray(wf.kys))[None, :] < 1.0)
    red = dict(
        spectrum=tac.spectrum(), spectrum0=tac.spectrum(0), diffraction=tac.diffraction(),
        spectral_diffraction=tac.spectral_diffraction(20.0), spectrum_image=tac.spectrum_image(20.0),
        dispersion=tac.dispersion(kx_path, ky_path))
    np.savez_compressed(os.path.join(HERE, "tacaw48.npz"), wf=wf.wavefunction_data.numpy(),
                        intensity=inten, frequencies=tac.frequencies, kx_path=kx_path, ky_path=ky_path,
                        mask=mask, **red)


def probe_kat():
    xs = np.linspace(0, 50, 501)
    ys = np.linspace(0, 49, 491)
    out = {}
    for mrad in (1, 3, 5, 15, 30):
        p = Probe(xs, ys, mrad=mrad, eV=100e3, device="cpu")
        out[f"mrad{mrad}"] = p.array.numpy()[::5, ::5].astype(np.complex64)
    np.savez_compressed(os.path.join(HERE, "probe_kat.npz"), **out)


def si_c1():
    traj = synthetic.silicon_trajectory(cells=(5, 5, 10), a=5.11, n_frames=1, seed=0)
    calc, wf = ref_run(traj, aperture=0.0, voltage_eV=100e3)
    pot = Potential(calc.xs, calc.ys, calc.zs, traj.positions[0], names_for(calc, traj.atom_types),
                    kind="kirkland", device="cpu").array.numpy()
    np.savez_compressed(
        os.path.join(HERE, "si_c1.npz"),
        wf=wf.wavefunction_data.numpy()[0, 0, :, :, 0].astype(np.complex64),
        pot_planes=pot[:, :, [0, 1, 51, 102]].astype(np.float32),
        pot_sum_z=pot.sum(axis=2), pot_sum_xy=pot.sum(axis=(0, 1)))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["small64", "tacaw48", "probe_kat", "si_c1"]
    for name in which:
        globals()[name]()
        print("wrote", name)
