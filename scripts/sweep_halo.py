"""Pipeline-depth / store-path sweep of the halo conv kernel and A/B of the weight-gradient kernels (tuning API).
   python scripts/sweep_halo.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K, conv as C, lib


def bench(fn, n_rot, iters=20):
    for i in range(3): fn(i % n_rot)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i % n_rot)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda"
for (H, W, I, O) in [(64, 256, 128, 128), (32, 128, 128, 128), (16, 64, 256, 256), (64, 256, 64, 64)]:
    g = C.plain_geom(H, W, I, O, 3)
    n_rot = min(32, max(2, int(300e6 // (B * H * W * I * 2)) + 1))
    xs = [torch.randn(B, H, W, I, device=dev).bfloat16() for _ in range(n_rot)]
    gys = [torch.randn(B, H, W, O, device=dev).bfloat16() for _ in range(n_rot)]
    w = (torch.randn(O, 9 * I, device=dev) / (9 * I) ** 0.5).bfloat16()
    out = torch.empty(B, H, W, O, device=dev, dtype=torch.bfloat16)
    epi = dict(col_scale=torch.rand(B, O, device=dev) + 0.5, noise=torch.randn(B, H, W, device=dev),
               noise_strength=torch.ones(1, device=dev), bias=torch.randn(O, device=dev), act=1, act_gain=1.4)
    fl = 2.0 * B * H * W * 9 * I * O
    res = []
    for a_st, b_st, staged in [(2, 4, 1), (2, 6, 0), (2, 4, 0), (3, 3, 0), (3, 4, 0), (3, 2, 1), (2, 8, 0)]:
        lib.set_tuning("halo_a_stages", a_st); lib.set_tuning("halo_b_stages", b_st); lib.set_tuning("halo_staged", staged)
        t = bench(lambda i: K.conv2d_igemm(xs[i], w, out=out, **g.kernel_kwargs(), **epi), n_rot)
        res.append(f"A{a_st}B{b_st}{'s' if staged else 'd'} {t:6.1f}us {fl / t / 1e6:5.0f}TF")
    print(f"conv {H}x{W} {I}->{O} B={B} | " + " | ".join(res), flush=True)
    lib.set_tuning("halo_a_stages", 2); lib.set_tuning("halo_b_stages", 4); lib.set_tuning("halo_staged", 1)
    gw = torch.zeros(O, 9 * I, device=dev)
    res = []
    for halo in (1, 0):
        lib.set_tuning("wgrad_halo", halo)
        t = bench(lambda i: K.conv2d_wgrad(xs[i], gys[i], gw=gw, **g.kernel_kwargs()), n_rot)
        res.append(f"{'halo' if halo else 'plain'} {t:6.1f}us {fl / t / 1e6:5.0f}TF")
    lib.set_tuning("wgrad_halo", 1)
    print(f"wgrad {H}x{W} {I}->{O} B={B} | " + " | ".join(res), flush=True)
