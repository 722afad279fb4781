"""HBM stream bandwidth by direction (torch library kernels, 4 GiB buffers): write-only, read-only, copy."""
import torch
n = 1 << 30   # float32 elements = 4 GiB
a = torch.empty(n, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
s = t(lambda: a.fill_(1.0)); print(f"write-only  (fill_)  {4*n/s/1e9:8.1f} GB/s")
s = t(lambda: a.sum());      print(f"read-only   (sum)    {4*n/s/1e9:8.1f} GB/s")
s = t(lambda: b.copy_(a));   print(f"copy        (copy_)  {8*n/s/1e9:8.1f} GB/s (read + write bytes)")
