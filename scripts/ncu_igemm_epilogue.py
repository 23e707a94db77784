"""Two launches each of epilogue-dominated conv_igemm shapes, for `ncu --set full --import-source on`:
1x1 64->128 on 128x32x128 (the residual skip GEMM), 3x3 stride-2 64->128 on 128x66x258 and the 3x3 64->64 on 128x64x256
with the halo kernel off (N = 64 tiles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K, lib

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def go(xshape, wshape, taps, pad, stride):
    B, H, W, I = xshape
    Ho = (H + 2 * pad[0] - taps[0]) // stride[0] + 1
    Wo = (W + 2 * pad[1] - taps[1]) // stride[1] + 1
    x = torch.randn(*xshape, device=dev).bfloat16()
    w = (torch.randn(*wshape, device=dev) / wshape[1] ** 0.5).bfloat16()
    for _ in range(2):
        flush.zero_()
        K.conv2d_igemm(x, w, Ho=Ho, Wo=Wo, taps=taps, pad=pad, stride=stride, up=(0, 0))


lib.set_tuning("conv_halo", 0)
go((128, 32, 128, 64), (128, 64), (1, 1), (0, 0), (1, 1))
go((128, 66, 258, 64), (128, 576), (3, 3), (0, 0), (2, 2))
go((128, 64, 256, 64), (64, 576), (3, 3), (1, 1), (1, 1))
torch.cuda.synchronize()
print("done")
