"""`HAADFData(wfdata).calculateADF()`: annular dark-field signal from the exit waves
(reference src/postprocessing/haadf_data.py:35-72), detector sum on the GPU."""
from __future__ import annotations

import numpy as np
import torch

from .. import engine
from .wf_data import WFData


class HAADFData(WFData):
    def __init__(self, WFData):
        self.__class__ = type(WFData.__class__.__name__, (self.__class__, WFData.__class__), {})
        self.__dict__ = WFData.__dict__

    def calculateADF(self, collection_angle: float = 45, preview: bool = False):
        """adf[i, j] = mean_frames sum_k |psi_k| * (|k| > collection_angle*1e-3/lambda) for the probe
        nearest to the i-th unique x / j-th unique y (haadf_data.py:43-65).  The mask uses the
        float32 kxs/kys labels of the WFData, like the reference.

        Frame-sharded runs (WFData.shard.world > 1: this process holds only its block of frames): the per-probe sums
        are combined over ranks with one all-reduce and divided by the GLOBAL frame count, so every rank returns the
        same image -- which makes this call a collective: every rank of the run must make it."""
        pp = np.asarray(self.probe_positions)
        self.xs = torch.as_tensor(sorted(set(pp[:, 0])))
        self.ys = torch.as_tensor(sorted(set(pp[:, 1])))
        kxs = torch.as_tensor(self.kxs)
        kys = torch.as_tensor(self.kys)
        q = torch.sqrt(kxs[:, None] ** 2 + kys[None, :] ** 2)
        radius = (collection_angle * 1e-3) / self.probe.wavelength
        mask = torch.zeros(q.shape)
        mask[q > radius] = 1
        from .wf_data import SlabStore
        shard = getattr(self, "shard", None)
        sums_dev = getattr(self, "adf_sums", None)
        if sums_dev is not None:
            # detector-only run (MultisliceCalculator.setup(adf_collection_angle=...)): the sums were taken after each exit FFT
            if abs(float(self.adf_collection_angle) - float(collection_angle)) > 1e-12:
                raise ValueError(f"this WFData holds detector sums for a collection angle of {self.adf_collection_angle} mrad only")
            sums = sums_dev[-1]                                                  # exit layer, (P, T_local)
            P, T = sums.shape
        elif isinstance(self.wavefunction_data, SlabStore):
            store = self.wavefunction_data
            dev = store.device
            P, T = store.P, store.T
            sums = torch.zeros((T * P,), dtype=torch.float64, device=dev)
            for h in range(store.world):                                         # detector rows of every block
                blk = store.block(h, store.L - 1)                                # (T, P, rows_h, ny)
                m = mask[store.row_starts[h]:store.row_starts[h] + store.row_counts[h]].to(dev, torch.float32).reshape(-1)
                sums += engine.sum_pixels(blk.reshape(T * P, -1), m.contiguous())
            sums = sums.reshape(T, P).t()
        else:
            wf = self.wavefunction_data[:, :, :, :, -1]
            if not hasattr(wf, "dim"):
                wf = torch.from_numpy(np.asarray(wf))
            dev = engine._device(wf.device if wf.device.type == "cuda" else None)
            wf = wf.to(device=dev, dtype=torch.complex64).contiguous()
            P, T, nx, ny = wf.shape
            sums = engine.sum_pixels(wf.reshape(P * T, nx * ny), mask.to(dev, torch.float32).reshape(-1)).reshape(P, T)
        if shard is not None and shard.world > 1:
            import torch.distributed as dist
            tot = sums.sum(dim=1)
            dist.all_reduce(tot)
            per_probe = (tot / float(shard.total)).cpu()
        else:
            per_probe = sums.mean(dim=1).cpu()
        self.adf = torch.zeros((len(self.xs), len(self.ys)))
        for i, x in enumerate(self.xs.tolist()):
            for j, y in enumerate(self.ys.tolist()):
                p = int(np.argmin(np.sqrt(np.sum((pp - np.array([x, y])[None, :]) ** 2, axis=1))))
                self.adf[i, j] = per_probe[p]
        return self.adf

    def plot(self):
        import matplotlib.pyplot as plt
        fig, ax = plt.subplots()
        ax.imshow(self.adf.T, cmap="inferno",
                  extent=(float(self.xs.min()), float(self.xs.max()), float(self.ys.min()), float(self.ys.max())))
        plt.show()
