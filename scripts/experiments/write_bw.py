"""Pure-write HBM bandwidth (torch fill of a 1 GB buffer) next to the copy bandwidth: the ceiling of write-only kernels
such as FromRGB forward (3 channels in, 64 out)."""
import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
tw = t(lambda: x.zero_())
tc = t(lambda: y.copy_(x))
print(f"fill 1 GiB: {tw * 1e6:.1f} us = {(1 << 30) / tw / 1e9:.0f} GB/s written")
print(f"copy 1 GiB: {tc * 1e6:.1f} us = {2 * (1 << 30) / tc / 1e9:.0f} GB/s read + written")
xs = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tw = t(lambda: xs.zero_())
print(f"fill 256 MiB: {tw * 1e6:.1f} us = {(256 << 20) / tw / 1e9:.0f} GB/s written")
