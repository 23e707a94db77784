"""Two launches each of generic conv_wgrad shapes for `ncu --set full --import-source on`: the stride-2 role-swapped
gradient of the 32x128 -> 64x256 up layer, the 8x32 256 -> 256 plain gradient, the 4x16 512 -> 512 one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K, lib

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def go(xshape, O, taps, pad, stride):
    B, H, W, I = xshape
    Ho = (H + 2 * pad[0] - taps[0]) // stride[0] + 1
    Wo = (W + 2 * pad[1] - taps[1]) // stride[1] + 1
    x = torch.randn(*xshape, device=dev).bfloat16()
    gy = torch.randn(B, Ho, Wo, O, device=dev).bfloat16()
    gw = torch.zeros(O, taps[0] * taps[1] * I, device=dev)
    for _ in range(2):
        flush.zero_()
        K.conv2d_wgrad(x, gy, gw=gw, Ho=Ho, Wo=Wo, taps=taps, pad=pad, stride=stride, up=(0, 0))


go((64, 66, 258, 128), 128, (3, 3), (0, 0), (2, 2))
go((64, 8, 32, 256), 256, (3, 3), (1, 1), (1, 1))
go((64, 4, 16, 512), 512, (3, 3), (1, 1), (1, 1))
torch.cuda.synchronize()
print("done")
