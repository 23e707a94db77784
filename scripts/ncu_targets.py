"""Launch the dominant tensor-core kernels of the training step a few times each, for `ncu --set full`:
    ncu --set full --clock-control none --import-source on -k regex:conv -c 12 -o gpurun_out/x python scripts/ncu_targets.py [batch]
Order of launches (2 each, L2 flushed in between): conv3x3_halo 64x256 128->128 (the dominant modulated-conv launch of
BASELINE configs[2], full epilogue) | conv3x3_halo 32x128 128->128 | conv_wgrad_halo 64x256 128->128 | conv_igemm stride-2
3x3 on 66x258 (the up layer's input gradient) | conv_wgrad (generic) 16x64 role-swapped shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K, conv as C

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def tensors(g):
    oh, ow = g.out_hw
    x = torch.randn(B, g.H, g.W, g.cin, device=dev).bfloat16()
    gy = torch.randn(B, oh, ow, g.cout, device=dev).bfloat16()
    w = (torch.randn(g.n_total, g.k_total, device=dev) / g.k_total ** 0.5).bfloat16()
    return x, gy, w


def epi(g):
    return dict(col_scale=torch.rand(B, g.cout, device=dev) + 0.5, noise=torch.randn(B, g.H, g.W, device=dev),
                noise_strength=torch.ones(1, device=dev), bias=torch.randn(g.cout, device=dev), act=1, act_gain=1.4)


g_top, g_mid = C.plain_geom(64, 256, 128, 128, 3), C.plain_geom(32, 128, 128, 128, 3)
for g in (g_top, g_mid):
    x, gy, w = tensors(g)
    e = epi(g)
    for _ in range(2):
        flush.zero_()
        K.conv2d_igemm(x, w, **g.kernel_kwargs(), **e)
    if g is g_top:
        for _ in range(2):
            flush.zero_()
            K.conv2d_wgrad(x, gy, **g.kernel_kwargs())
    del x, gy, w
spec = C.weight_spec("upT", 32, 128, 128, 128, 3, True, "modconv")
gT = torch.randn(B, 66, 258, 128, device=dev).bfloat16()
wa = (torch.randn(spec.adj_rows, spec.adj_cols, device=dev) / 34.0).bfloat16()
for _ in range(2):
    flush.zero_()
    K.conv2d_igemm(gT, wa, **spec.s2_kwargs)
xs = torch.randn(B, 32, 128, 128, device=dev).bfloat16()
for _ in range(2):
    flush.zero_()
    K.conv2d_wgrad(gT, xs, **spec.s2_kwargs)
torch.cuda.synchronize()
print("done")
