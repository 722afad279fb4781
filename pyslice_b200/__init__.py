"""pyslice_b200 -- B200-native (sm_100a) multislice + TACAW engine behind PySlice's Python API.

Import names mirror the reference (h-walk/PySlice) so callers only change the package prefix:

    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.multislice.multislice import Probe, probe_grid, create_batched_probes, Propagate
    from pyslice_b200.multislice.potentials import gridFromTrajectory, Potential
    from pyslice_b200.multislice.trajectory import Trajectory
    from pyslice_b200.postprocessing.wf_data import WFData
    from pyslice_b200.postprocessing.tacaw_data import TACAWData
    from pyslice_b200.postprocessing.haadf_data import HAADFData

All arithmetic on the hot path runs in hand-written CUDA kernels (`csrc/`, one C-ABI shared
library `libpsb.so`, declared in `include/pyslice_b200.h`) called through ctypes; torch tensors
are only device-memory containers.  There is no CPU fallback: using the engine without the
built library or without a CUDA device raises.
"""
__version__ = "0.1.0"
