"""Launch the dominant tensor-core kernels of the training step a few times each, for `ncu --set full`:
    ncu --set full --clock-control none --import-source on -k regex:conv_ -c 12 -o gpurun_out/x python scripts/ncu_targets.py
Order of launches (3 each): igemm fwd 32x128 128->128 | igemm fwd 16x64 256->256 | wgrad 16x64 256->256 | wgrad 32x128 128->128
(config-1 shapes, batch 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K, conv as C

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"
def tensors(g):
    oh, ow = g.out_hw
    x = torch.randn(B, g.H, g.W, g.cin, device=dev).bfloat16()
    gy = torch.randn(B, oh, ow, g.cout, device=dev).bfloat16()
    w = (torch.randn(g.n_total, g.k_total, device=dev) / g.k_total ** 0.5).bfloat16()
    return x, gy, w
g1, g2 = C.plain_geom(32, 128, 128, 128, 3), C.plain_geom(16, 64, 256, 256, 3)
t1, t2 = tensors(g1), tensors(g2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for g, (x, gy, w) in ((g1, t1), (g2, t2)):
    for _ in range(3):
        flush.zero_()
        K.conv2d_igemm(x, w, **g.kernel_kwargs())
for g, (x, gy, w) in ((g2, t2), (g1, t1)):
    for _ in range(3):
        flush.zero_()
        K.conv2d_wgrad(x, gy, **g.kernel_kwargs())
torch.cuda.synchronize()
print("done")
