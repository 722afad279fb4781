"""ctypes binding of libpsb.so (the C ABI declared in include/pyslice_b200.h).

The library is built in-tree (`make -C pyslice_b200/csrc` or `__graft_entry__.build()`).  There is
no fallback: if it is missing, or no CUDA device is present, using the engine raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpsb.so")

_lib = None
_emulated = False      # set only by tests/emu/ (kernel emulator for the CPU test suite)


class PsbError(RuntimeError):
    """A libpsb call returned a negative status."""


_P = C.c_void_p
_I = C.c_int
_LL = C.c_longlong
_F = C.c_float
_D = C.c_double

class PropagateDesc(C.Structure):
    """`psb_propagate_desc` of include/pyslice_b200.h (field for field)."""
    _fields_ = [
        ("struct_bytes", _I), ("mode", _I), ("layer_every", _I),
        ("n_frames", _I), ("n_probes", _I), ("nz", _I), ("nx", _I), ("ny", _I),
        ("probes", _P), ("t", _P), ("phase", _P), ("t0_scratch", _P), ("prop_x", _P), ("prop_y", _P), ("psi_work", _P),
        ("wf_out", _P), ("stride_probe", _LL), ("stride_frame", _LL), ("stride_layer", _LL),
        ("slab_world", _I), ("slab_layers", _I), ("slab_frames", _I), ("slab_probes", _I), ("frame0", _I), ("probe0", _I),
        ("det_mask", _P), ("det_out", _P), ("det_stride_layer", _LL), ("det_stride_probe", _LL), ("det_scratch", _P),
        ("stream", _P),
    ]


_SIGNATURES = {
    "psb_version": (C.c_int, []),
    "psb_last_error": (C.c_char_p, []),
    "psb_sm_count": (C.c_int, []),
    "psb_release_tables": (None, []),
    "psb_launch_count": (C.c_longlong, []),
    "psb_set_fast_path": (None, [_I]),
    "psb_set_graph_mode": (None, [_I]),
    "psb_set_sf_mode": (None, [_I]),
    "psb_bin_atoms": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P, _D, _D, _D, _P, _P, _P, _P, _P, _P]),
    "psb_build_transmission": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _F, _F, _P, _P, _P, _LL, _P]),
    "psb_transmission_from_potential": (C.c_int, [_P, _P, _LL, _F, _P]),
    "psb_fft2": (C.c_int, [_P, _P, _I, _I, _I, _I, _F, _P]),
    "psb_shift_probes": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "psb_build_phase": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _F, _F, _P, _P, _LL, _P]),
    "psb_phase_format_supported": (C.c_int, [_I, _I]),
    "psb_propagate_phase": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _LL, _LL, _LL, _I, _P]),
    "psb_propagate": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _LL, _LL, _LL, _I, _P]),
    "psb_propagate_ex": (C.c_int, [C.POINTER(PropagateDesc)]),
    "psb_tacaw_intensity": (C.c_int, [_P, _LL, _LL, _I, _I, _LL, _P, _P]),
    "psb_sum_pixels": (C.c_int, [_P, _P, _I, _LL, _LL, _P, _P]),
    "psb_sum_abs_pixels": (C.c_int, [_P, _P, _I, _LL, _LL, _P, _P]),
    "psb_sum_frames": (C.c_int, [_P, _I, _I, _LL, _P, _P]),
}

EXPORTS = tuple(_SIGNATURES)


def _bind(handle):
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)      # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return handle


def load(path: str = LIB_PATH):
    """dlopen the library and declare every prototype (no CUDA call is made)."""
    if not os.path.exists(path):
        raise PsbError(
            f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
            "or make -C pyslice_b200/csrc). pyslice_b200 has no CPU fallback.")
    return _bind(C.CDLL(path))


def lib():
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def check(rc: int, what: str = "libpsb"):
    if rc != 0:
        msg = lib().psb_last_error()
        raise PsbError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def is_emulated() -> bool:
    return _emulated


def _use_emulator(path: str):
    """TEST HOOK (tests/emu only): bind the g++-compiled kernel emulator instead of libpsb.so."""
    global _lib, _emulated
    _lib = _bind(C.CDLL(path))
    _emulated = True


def _reset():
    global _lib, _emulated
    _lib = None
    _emulated = False
