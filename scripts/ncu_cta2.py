import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K, lib
dev = "cuda"
B, H, W, I, O = 64, 64, 256, 128, 128
x = torch.randn(B, H, W, I, device=dev).bfloat16()
w = (torch.randn(O, 9 * I, device=dev) / (9 * I) ** 0.5).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for v in (1, 0):
    lib.set_tuning("halo_cta2", v)
    for _ in range(2):
        flush.zero_()
        K.conv2d_igemm(x, w, Ho=H, Wo=W, taps=(3, 3), pad=(1, 1), stride=(1, 1), up=(0, 0))
torch.cuda.synchronize()
print("done")
