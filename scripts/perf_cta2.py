"""conv3x3_halo_kernel: single CTA vs CTA pairs (tcgen05 cta_group::2) on the halo-kernel shapes of configs[2], isolated.
   python scripts/perf_cta2.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K, lib


def bench(fn, n_rot, iters=20):
    for i in range(3): fn(i % n_rot)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i % n_rot)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
for (H, W, I, O, up) in [(64, 256, 128, 128, 0), (32, 128, 128, 128, 0), (16, 64, 256, 256, 0), (32, 128, 128, 128, 1),
                         (16, 64, 256, 128, 1), (64, 256, 64, 64, 0)]:
    nph = 4 if up else 1
    n_rot = min(32, max(2, int(300e6 // (B * H * W * I * 2)) + 1))
    xs = [torch.randn(B, H, W, I, device=dev).bfloat16() for _ in range(n_rot)]
    w = (torch.randn(nph * O, 9 * I, device=dev) / (9 * I) ** 0.5).bfloat16()
    oh, ow = H * (2 if up else 1), W * (2 if up else 1)
    out = torch.empty(B, oh, ow, O, device=dev, dtype=torch.bfloat16)
    epi = dict(col_scale=torch.rand(B, O, device=dev) + 0.5, noise=torch.randn(B, oh, ow, device=dev),
               noise_strength=torch.ones(1, device=dev), bias=torch.randn(O, device=dev), act=1, act_gain=1.4)
    kw = dict(Ho=H, Wo=W, taps=(3, 3), pad=(1, 1), stride=(1, 1), up=(up, up))
    fl = 2.0 * B * H * W * 9 * I * O * nph
    res = []
    for cta2, ast, bst, stg in [(0, 2, 4, 1), (0, 3, 3, 0), (1, 2, 4, 1), (1, 3, 2, 0), (1, 3, 3, 0), (1, 2, 4, 0)]:
        lib.set_tuning("halo_cta2", cta2); lib.set_tuning("halo_a_stages", ast); lib.set_tuning("halo_b_stages", bst)
        lib.set_tuning("halo_staged", stg)
        t = bench(lambda i: K.conv2d_igemm(xs[i], w, out=out, **kw, **epi), n_rot)
        res.append(f"cta2={cta2} A{ast}B{bst}{'s' if stg else 'd'} {t:6.1f}us {fl / t / 1e6:5.0f}TF")
    lib.set_tuning("halo_cta2", 0); lib.set_tuning("halo_a_stages", 2); lib.set_tuning("halo_b_stages", 4); lib.set_tuning("halo_staged", 1)
    print(f"conv {H}x{W} {I}->{O}{' up' if up else ''} B={B} | " + " | ".join(res), flush=True)
