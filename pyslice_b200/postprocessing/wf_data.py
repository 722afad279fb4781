"""`WFData`: the output container (reference src/postprocessing/wf_data.py:9-28)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Tuple

import numpy as np


@dataclass
class WFData:
    """probe_positions, time (ps), kxs, kys, layer, wavefunction_data (probe, time, kx, ky, layer), probe.

    `wavefunction_data[p, f, :, :, l]` is the fftshifted, unnormalised forward FFT of the wave
    function of probe p / frame f after layer l (reciprocal space), a complex64 CUDA tensor here.

    Multi-GPU runs add `shard` (calculators.FrameShard): `wavefunction_data` then holds only this rank's block of frames
    (`shard.counts[shard.rank]` of them, starting at `shard.start`) while `time` describes the whole run.  TACAWData and
    HAADFData know about it (their reducers are collectives then: call them on every rank); code that reads
    `wavefunction_data` directly must index frames relative to `shard.start`.
    """
    probe_positions: List[Tuple[float, float]]
    time: np.ndarray
    kxs: Any
    kys: Any
    layer: np.ndarray
    wavefunction_data: Any
    probe: Any
