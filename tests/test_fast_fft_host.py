"""CPU check of the fused slice-step kernels' line transform (pyslice_b200/csrc/fast_fft.cuh): the same
source compiled for the host (scalar arithmetic instead of packed fp32x2 PTX), one host thread per CUDA
thread of a line, against numpy.  Covers stage index arithmetic, per-thread twiddle selection and the
alternating exchange buffers; the PTX paths are covered by the -m gpu suite."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "emu", "libfastfft_emu.so")


@pytest.fixture(scope="module")
def harness():
    subprocess.run(["make", "-C", os.path.join(ROOT, "pyslice_b200", "csrc"), "emu", "-j8"], check=True,
                   stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(LIB)
    lib.fast_fft_line.restype = ctypes.c_int
    lib.fast_fft_line.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return lib


@pytest.mark.parametrize("N", [256, 512, 1024, 2048])
@pytest.mark.parametrize("direction", [-1, 1])
def test_line_fft_matches_numpy(harness, N, direction):
    rng = np.random.default_rng(N + direction)
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    out = np.empty(N, np.complex64)
    assert harness.fast_fft_line(N, direction, x.ctypes.data, out.ctypes.data, 1) == 0
    ref = np.fft.fft(x.astype(np.complex128)) if direction < 0 else np.fft.ifft(x.astype(np.complex128)) * N
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 5e-7
    # impulse response: every output bin is exactly a unit phasor (twiddle table entries used as is)
    imp = np.zeros(N, np.complex64)
    imp[3] = 1
    assert harness.fast_fft_line(N, direction, imp.ctypes.data, out.ctypes.data, 1) == 0
    k = np.arange(N)
    want = np.exp(direction * 2j * np.pi * 3 * k / N)
    assert np.abs(out - want).max() < 5e-7


@pytest.mark.parametrize("N", [256, 512, 1024, 2048])
def test_back_to_back_transforms_share_exchange_buffers(harness, N):
    """two transforms in a row (as in a tile: FFT a, multiply, FFT b) reuse the alternating buffers safely"""
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    out = np.empty(N, np.complex64)
    assert harness.fast_fft_line(N, -1, x.ctypes.data, out.ctypes.data, 2) == 0
    ref = np.fft.fft(np.fft.fft(x.astype(np.complex128)))
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-6


def test_unsupported_size_is_rejected(harness):
    x = np.zeros(128, np.complex64)
    assert harness.fast_fft_line(128, -1, x.ctypes.data, x.ctypes.data, 1) == -1
