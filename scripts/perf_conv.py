"""Isolated timings of the 3x3 convolution kernels per layer shape (rotating inputs > L2, CUDA events):
conv_igemm_kernel with staged / direct epilogue stores and the halo-reuse kernel, bare and with the full
modulated-conv epilogue.   python scripts/perf_conv.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textboxgan_b200 import kernels as K, conv as C, lib


def bench(fn, n_rot, iters=30):
    for i in range(3): fn(i % n_rot)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i % n_rot)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def run(B, H, W, I, O):
    dev = "cuda"
    g = C.plain_geom(H, W, I, O, 3)
    n_rot = min(48, max(2, int(300e6 // (B * H * W * I * 2)) + 1))
    xs = [torch.randn(B, H, W, I, device=dev).bfloat16() for _ in range(n_rot)]
    w = (torch.randn(O, 9 * I, device=dev) / (9 * I) ** 0.5).bfloat16()
    out = torch.empty(B, H, W, O, device=dev, dtype=torch.bfloat16)
    epi = dict(col_scale=torch.rand(B, O, device=dev) + 0.5, noise=torch.randn(B, H, W, device=dev),
               noise_strength=torch.ones(1, device=dev), bias=torch.randn(O, device=dev), act=1, act_gain=1.4)
    fl = 2.0 * B * H * W * 9 * I * O
    cols = []
    for name, tune in (("igemm", dict(conv_halo=0)),
                       ("halo", dict(conv_halo=1))):
        for k, v in tune.items():
            lib.set_tuning(k, v)
        tb = bench(lambda i: K.conv2d_igemm(xs[i], w, out=out, **g.kernel_kwargs()), n_rot)
        tf = bench(lambda i: K.conv2d_igemm(xs[i], w, out=out, **g.kernel_kwargs(), **epi), n_rot)
        cols.append(f"{name}: bare {tb:6.1f} us {fl / tb / 1e6:6.0f} TF/s, full-epilogue {tf:6.1f} us {fl / tf / 1e6:6.0f} TF/s")
    lib.set_tuning("conv_halo", 1)
    print(f"{H}x{W} {I}->{O} B={B} | " + " | ".join(cols), flush=True)


B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for shp in [(16, 64, 64, 64), (16, 64, 256, 256), (32, 128, 128, 128), (64, 256, 128, 128), (64, 256, 64, 64), (16, 64, 128, 128)]:
    run(B, *shp)
