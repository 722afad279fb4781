// placeholder for the fused persistent slice-step kernels (filled in by the optimisation milestone)
