"""Host-side logic that needs neither GPU nor emulator."""
import numpy as np
import pytest

from oracle import pyslice_oracle as orc
from pyslice_b200 import hostmath
from pyslice_b200.multislice.calculators import split_frames
from pyslice_b200.multislice.multislice import probe_grid
from pyslice_b200.multislice.trajectory import Trajectory
from pyslice_b200 import synthetic


def test_hostmath_matches_oracle_tables():
    assert hostmath.wavelength(100e3) == orc.wavelength(100e3)
    assert hostmath.interaction_sigma(60e3) == orc.interaction_sigma(60e3)
    box = np.diag([25.55, 27.15, 51.1])
    a = hostmath.grid_from_box(box)
    b = orc.grid_from_box(box)
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
    lo, hi, dz = hostmath.slice_bounds(a[2])
    lo2, hi2 = orc.slice_bounds(a[2])
    assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
    kxs, kys = hostmath.kgrid(a[0], a[1])
    ff = hostmath.form_factor_table(kxs, kys, [5, 14])
    qsq = kxs[:, None] ** 2 + kys[None, :] ** 2
    assert np.allclose(ff[1], orc.form_factor(qsq, 14), rtol=1e-14)
    px, py = hostmath.propagator_tables(kxs, kys, hostmath.wavelength(100e3), dz)
    P = orc.fresnel_propagator(a[0], a[1], a[2], 100e3)
    assert np.allclose(px[:, None] * py[None, :], P, rtol=0, atol=1e-13)


def test_slice_bounds_have_gaps_and_overlaps():
    # the reference's expressions leave 1-ulp gaps / overlaps between neighbouring slices
    zs = np.linspace(0, 305.3, int(305.3 / 0.5) + 1, endpoint=False)
    lo, hi, _ = hostmath.slice_bounds(zs)
    assert (hi[:-1] != lo[1:]).any()


def test_split_frames_ragged():
    assert split_frames(100, 8) == [13, 13, 13, 13, 12, 12, 12, 12]
    assert sum(split_frames(7, 2)) == 7 and split_frames(2, 4) == [1, 1, 0, 0]


def test_probe_grid_matches_reference_layout():
    xy = probe_grid([0, 1], [0, 2], 3, 2)
    assert xy.shape == (6, 2)
    assert np.allclose(xy[:, 0], [0, .5, 1, 0, .5, 1]) and np.allclose(xy[:, 1], [0, 0, 0, 2, 2, 2])


def test_trajectory_validation_errors():
    pos = np.zeros((2, 3, 3))
    with pytest.raises(ValueError, match="positions must be"):
        Trajectory(np.zeros(3), np.zeros((2, 3)), pos, np.eye(3), 0.01)
    with pytest.raises(ValueError, match="Atom count mismatch"):
        Trajectory(np.zeros(4), pos, pos, np.eye(3), 0.01)
    with pytest.raises(ValueError, match="box_matrix"):
        Trajectory(np.zeros(3), pos, pos, np.eye(2), 0.01)
    t = Trajectory(np.zeros(3), pos, pos, np.eye(3), 0.01)
    assert t.n_frames == 2 and t.n_atoms == 3
    with pytest.raises(ValueError):
        t.slice_timesteps([])
    assert t.tile_positions((2, 1, 1)).n_atoms == 6
    assert t.slice_timesteps([1]).n_frames == 1


def test_synthetic_grids_hit_requested_sizes():
    for cells, a, want in [((5, 5, 10), 5.11, (256, 256, 103)), ((5, 5, 50), 5.11, (256, 256, 512)),
                           ((20, 20, 12), 5.1175, (1024, 1024, 123))]:
        box = np.diag([cells[0] * a, cells[1] * a, cells[2] * a])
        xs, ys, zs, *_ = hostmath.grid_from_box(box)
        assert (len(xs), len(ys), len(zs)) == want
    t = synthetic.hbn_graphene_trajectory(n_frames=1)
    xs, ys, zs, *_ = hostmath.grid_from_box(t.box_matrix)
    assert (len(xs), len(ys), len(zs)) == (512, 512, 67) and t.n_atoms == 9600


def test_batch_and_chunk_sizing_rules(monkeypatch):
    """engine._fewest_rounds / chunk_images: host-side sizing against the persistent grids (296 CTA slots on 148 SMs)"""
    from types import SimpleNamespace as NS

    from pyslice_b200 import engine
    monkeypatch.setattr(engine, "_sm_count", lambda: 148)
    # C2: 500 frames of one 256 x 256 probe (16 column tiles per image): 127 per batch = 3 x 7 + 7 rounds
    fb = engine._fewest_rounds(500, 127, 16, 296)
    rounds = lambda fb: (500 // fb) * -(-fb * 16 // 296) + (-(-(500 % fb) * 16 // 296) if 500 % fb else 0)
    assert rounds(fb) == 28 and rounds(100) == 30 and fb <= 127
    assert engine._fewest_rounds(3, 3, 16, 296) == 3 and engine._fewest_rounds(100, 1, 64, 296) == 1
    # potential chunks: as many images as the scratch holds, nudged only on large grids
    assert engine.chunk_images(NS(nx=256, ny=256, nz=512), 100) == 128
    assert engine.chunk_images(NS(nx=512, ny=512, nz=67), 100) == 32
    assert engine.chunk_images(NS(nx=1024, ny=1024, nz=123), 100) == 9          # 1152 tiles = 3.9 rounds, not 8 -> 3.5 of 4
    assert engine.chunk_images(NS(nx=256, ny=256, nz=9), 2) == 10                # never more than the work there is
    assert engine.chunk_images(NS(nx=4096, ny=4096, nz=40), 1) == 1
    # psi batches: L2-sized at 256 / 512 points, streaming-sized (up to 320 MB = 40 images, whole rounds of the one-CTA-per-SM
    # column pass preferred) at 1024 points
    cpu = NS(type="cpu")
    fb, pb = engine.batch_sizes(NS(nx=256, ny=256, nz=512, device=cpu), 1, 500)
    assert pb == 1 and 100 <= fb <= 160 and fb * 256 * 256 * 8 <= engine.PSI_BATCH_BYTES
    fb, pb = engine.batch_sizes(NS(nx=512, ny=512, nz=67, device=cpu), 256, 100)
    assert fb == 1 and pb * 512 * 512 * 8 <= engine.PSI_BATCH_BYTES and pb >= 32
    fb, pb = engine.batch_sizes(NS(nx=1024, ny=1024, nz=123, device=cpu), 1, 250)
    assert pb == 1 and 24 <= fb <= 40 and fb * 1024 * 1024 * 8 <= engine.PSI_BATCH_BYTES_STREAMING
