"""The C-ABI library loads and exports every symbol include/pyslice_b200.h declares (no compute)."""
import os
import re

from pyslice_b200 import _lib
from tests.helpers import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "pyslice_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_bound():
    declared = _declared()
    assert declared, "no prototypes found in the header"
    assert sorted(_lib.EXPORTS) == declared


def test_library_exports_every_symbol():
    # built by __graft_entry__.build(); dlopen + symbol lookup only, safe without a GPU
    handle = _lib.load()
    for name in _declared():
        assert hasattr(handle, name), name
    assert handle.psb_version() >= 100
    assert handle.psb_last_error() is not None


def test_missing_library_fails_loudly(tmp_path):
    import pytest
    with pytest.raises(_lib.PsbError):
        _lib.load(str(tmp_path / "nope.so"))


def test_no_cuda_no_fallback():
    import pytest
    import torch

    from pyslice_b200 import engine
    if torch.cuda.is_available() or _lib.is_emulated():
        pytest.skip("only meaningful on a CPU-only box with the real library")
    with pytest.raises(RuntimeError):
        engine._device(None)
