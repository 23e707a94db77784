"""Final-state captures for `ncu --set full`: the 1x1 skip GEMM and the stride-2 3x3 on conv_igemm (templated 8-warp
epilogue), fir4 (cp.async ring), bias_act_bwd<2> (wave-aware grid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def conv(xshape, wshape, taps, pad, stride):
    B, H, W, I = xshape
    Ho = (H + 2 * pad[0] - taps[0]) // stride[0] + 1
    Wo = (W + 2 * pad[1] - taps[1]) // stride[1] + 1
    x = torch.randn(*xshape, device=dev).bfloat16()
    w = (torch.randn(*wshape, device=dev) / wshape[1] ** 0.5).bfloat16()
    for _ in range(2):
        flush.zero_()
        K.conv2d_igemm(x, w, Ho=Ho, Wo=Wo, taps=taps, pad=pad, stride=stride, up=(0, 0))


conv((128, 32, 128, 64), (128, 64), (1, 1), (0, 0), (1, 1))
conv((128, 66, 258, 64), (128, 576), (3, 3), (0, 0), (2, 2))
x = torch.randn(64, 64, 256, 128, device=dev).bfloat16()
o = torch.randn(64, 64, 256, 128, device=dev).bfloat16()
for _ in range(2):
    flush.zero_()
    K.fir4(x, (66, 258), (-2, -2), 1.0 / 16.0)
nz = torch.randn(64, 64, 256, device=dev)
d = torch.rand(64, 128, device=dev) + 0.5
for _ in range(2):
    flush.zero_()
    K.bias_act_bwd(x, o, noise=nz, d=d, act=True, gain=1.4)
torch.cuda.synchronize()
print("done")
