"""Kernel emulator for the CPU test suite.

`libpsb_emu.so` is the engine's own kernel source (pyslice_b200/csrc) compiled by g++ with
-DPSB_EMU: every CUDA thread of a block becomes a host thread, __syncthreads a std::barrier.
It lets `pytest -m "not gpu"` run the real index arithmetic of the kernels on tiny problems and
compare it with the oracle before any GPU time is spent.  It is test infrastructure: the package
never loads it (pyslice_b200/_lib.py only binds libpsb.so) and it is far too slow to be a fallback.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libpsb_emu.so")


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "pyslice_b200", "csrc"), "emu", "-j8"], check=True,
                   stdout=subprocess.DEVNULL)
    return LIB


def activate():
    """Point pyslice_b200._lib at the emulator (tests only)."""
    from pyslice_b200 import _lib
    build()
    _lib._use_emulator(LIB)
