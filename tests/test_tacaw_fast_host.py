"""CPU check of the tiled mixed-radix TACAW time transform (pyslice_b200/csrc/tacaw_stages.cuh, the code
tacaw_fast_kernel runs between its barriers): the same source compiled for the host, one loop iteration per CUDA thread,
against numpy.  Covers the factorisation, every radix (2, 3, 4, 5), single-stage transforms, stage index arithmetic,
twiddle selection, the digit-reversal + fftshift permutation, ragged last tiles and strided frames; the launch
configuration and the barriers themselves are covered by the -m gpu suite."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "emu", "libtacaw_emu.so")


@pytest.fixture(scope="module")
def harness():
    subprocess.run(["make", "-C", os.path.join(ROOT, "pyslice_b200", "csrc"), "emu", "-j8"], check=True,
                   stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(LIB)
    lib.tacaw_fast_host.restype = ctypes.c_int
    lib.tacaw_fast_host.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
    lib.tacaw_fast_plan.restype = ctypes.c_int
    lib.tacaw_fast_plan.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def reference(x):
    """|fftshift_t FFT_t(psi - mean_t psi)|^2 (reference src/postprocessing/tacaw_data.py:89-104), float64"""
    x = x.astype(np.complex128)
    return np.abs(np.fft.fftshift(np.fft.fft(x - x.mean(axis=0, keepdims=True), axis=0), axes=0)) ** 2


@pytest.mark.parametrize("T,npix,want_px", [(20, 130, 64), (100, 70, 64), (60, 40, 64), (200, 21, 32), (250, 17, 32), (1000, 9, 8), (500, 19, 16), (2000, 9, 8), (4000, 5, 4),
                                            (64, 33, 64), (45, 7, 64), (6, 3, 64), (2, 5, 64), (3, 4, 64), (5, 65, 64),
                                            (1500, 8, 8), (1600, 3, 8)])
def test_tile_transform_matches_numpy(harness, T, npix, want_px):
    rng = np.random.default_rng(T)
    stride = npix + 3                                   # frames are not contiguous
    buf = (rng.normal(size=(T, stride)) + 1j * rng.normal(size=(T, stride)) + 2.0 - 1.0j).astype(np.complex64)
    out = np.full((T, npix), -1.0, np.float32)
    px = harness.tacaw_fast_host(T, npix, buf.ctypes.data, stride, out.ctypes.data)
    assert px == want_px
    ref = reference(buf[:, :npix])
    dc = T // 2
    keep = [i for i in range(T) if i != dc]
    assert np.abs(out[keep] - ref[keep]).max() <= 2e-5 * ref.max()
    assert np.all(out[dc] == 0.0)                       # psi - <psi> is exactly zero in the zero-frequency bin


@pytest.mark.parametrize("T,fac", [(20, [20]), (100, [10, 10]), (500, [10, 10, 5]), (2000, [20, 10, 10]),
                                   (4000, [20, 20, 10]), (50, [10, 5]), (40, [10, 4]), (48, [4, 4, 3]), (2, [2])])
def test_plan_factorisation_and_permutation(harness, T, fac):
    f = np.zeros(16, np.int32)
    perm = np.zeros(T, np.int32)
    n = harness.tacaw_fast_plan(T, f.ctypes.data, perm.ctypes.data)
    assert n == len(fac) and f[:n].tolist() == fac
    assert sorted(perm.tolist()) == list(range(T))      # a permutation
    assert perm[0] == T // 2                            # position 0 holds X[0], which fftshift puts at T // 2


def test_unsupported_lengths_are_refused(harness):
    f = np.zeros(16, np.int32)
    perm = np.zeros(97, np.int32)
    assert harness.tacaw_fast_plan(97, f.ctypes.data, perm.ctypes.data) == -1       # prime
    assert harness.tacaw_fast_plan(14, f.ctypes.data, perm.ctypes.data) == -1       # factor 7
    x = np.zeros((8000, 4), np.complex64)
    out = np.zeros((8000, 4), np.float32)
    assert harness.tacaw_fast_host(8000, 4, x.ctypes.data, 4, out.ctypes.data) == -1   # no tile of >= 4 pixels fits
