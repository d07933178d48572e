"""Probe: write-only, read-only and copy HBM bandwidth on this GPU (torch ops, CUDA events).  The step kernels'
traffic is write-dominated (outputs), so the copy peak in MEASURED_PEAKS.json is not the whole story."""
import torch

dev = "cuda:0"
n = 1 << 30   # 1 GiB of bytes
x = torch.empty(n, dtype=torch.uint8, device=dev)
y = torch.empty(n, dtype=torch.uint8, device=dev)
xf = x.view(torch.float32)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


t = timeit(lambda: x.zero_())
print(f"write-only (memset 1 GiB): {n / t / 1e9:.0f} GB/s")
t = timeit(lambda: xf.fill_(1.5))
print(f"write-only (fill kernel):  {n / t / 1e9:.0f} GB/s")
t = timeit(lambda: xf.sum())
print(f"read-only  (sum 1 GiB):    {n / t / 1e9:.0f} GB/s")
t = timeit(lambda: y.copy_(x))
print(f"copy (1 GiB -> 1 GiB):     {2 * n / t / 1e9:.0f} GB/s (read + write)")
