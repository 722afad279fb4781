"""Run a script (e.g. bench.py) against an alternative build of libpsb and/or with the chunk rule off (tuning only).
usage: PSB_VARIANT_LIB=pyslice_b200/libpsb_x.so PSB_CHUNK_RULE=0 python tools/run_variant.py bench.py --steps 2 ..."""
import os, runpy, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyslice_b200 import _lib, engine
if os.environ.get("PSB_VARIANT_LIB"):
    _lib._lib = _lib.load(os.path.abspath(os.environ["PSB_VARIANT_LIB"]))
if os.environ.get("PSB_SCRATCH_MB"):
    engine.SCRATCH_BYTES = int(os.environ["PSB_SCRATCH_MB"]) << 20
sys.argv = sys.argv[1:]
if sys.argv[0] == "-m":                      # python tools/run_variant.py -m pytest tests ...
    sys.argv = sys.argv[1:]
    runpy.run_module(sys.argv[0], run_name="__main__", alter_sys=True)
else:
    runpy.run_path(sys.argv[0], run_name="__main__")
