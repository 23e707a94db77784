"""Isolated timings (rotating inputs > L2, CUDA events, nothing else on the GPU) of the strided and 1x1 conv_igemm launch
shapes of a configs[2] iteration, with the tuning switches that apply to them:   python scripts/perf_shapes.py"""
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


def run(xshape, wshape, taps, pad, stride, tunes, wgrad_gy=None):
    dev = "cuda"
    B, H, W, I = xshape
    Ho = (H + 2 * pad[0] - taps[0]) // stride[0] + 1
    Wo = (W + 2 * pad[1] - taps[1]) // stride[1] + 1
    n_rot = min(32, max(2, int(300e6 // (B * H * W * I * 2)) + 1))
    xs = [torch.randn(B, H, W, I, device=dev).bfloat16() for _ in range(n_rot)]
    kw = dict(Ho=Ho, Wo=Wo, taps=taps, pad=pad, stride=stride, up=(0, 0))
    cols = []
    if wgrad_gy is None:
        O = wshape[0]
        w = (torch.randn(*wshape, device=dev) / wshape[1] ** 0.5).bfloat16()
        out = torch.empty(B, Ho, Wo, O, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * B * Ho * Wo * wshape[0] * wshape[1]
        byt = (B * H * W * I + B * Ho * Wo * O) * 2 / (stride[0] * stride[1] if taps == (1, 1) else 1)
        fn = lambda i: K.conv2d_igemm(xs[i], w, out=out, **kw)
    else:
        O = wgrad_gy
        gy = torch.randn(B, Ho, Wo, O, device=dev).bfloat16()
        gw = torch.zeros(O, taps[0] * taps[1] * I, device=dev)
        fl = 2.0 * B * Ho * Wo * O * taps[0] * taps[1] * I
        byt = (B * H * W * I + B * Ho * Wo * O) * 2
        fn = lambda i: K.conv2d_wgrad(xs[i], gy, gw=gw, **kw)
    for key, vals in tunes:
        for v in vals:
            lib.set_tuning(key, v)
            t = bench(fn, n_rot)
            cols.append(f"{key}={v}: {t:7.1f} us {fl / t / 1e6:7.1f} TF/s {byt / t / 1e3:6.0f} GB/s")
        lib.set_tuning(key, 1)
    kind = "wgrad" if wgrad_gy is not None else "conv "
    print(f"{kind} x{xshape} w{wshape if wgrad_gy is None else (O,)} t{taps} p{pad} s{stride} | " + " | ".join(cols), flush=True)


S2 = [("conv_halo", (1,))]
ST = [("conv_halo", (1,))]
run((128, 66, 258, 64), (128, 576), (3, 3), (0, 0), (2, 2), S2)
run((64, 66, 258, 128), (128, 1152), (3, 3), (0, 0), (2, 2), S2)
run((128, 34, 130, 128), (128, 1152), (3, 3), (0, 0), (2, 2), S2)
run((64, 34, 130, 128), (256, 1152), (3, 3), (0, 0), (2, 2), S2)
run((128, 10, 34, 256), (256, 2304), (3, 3), (0, 0), (1, 2), S2)
run((128, 32, 128, 64), (128, 64), (1, 1), (0, 0), (1, 1), ST)
run((128, 32, 128, 128), (64, 128), (1, 1), (0, 0), (1, 1), ST)
run((64, 16, 64, 64), (64, 64), (1, 1), (0, 0), (1, 1), ST)
run((64, 4, 32, 128), (128, 128), (1, 1), (0, 0), (1, 1), ST)
run((64, 2, 32, 256), (256, 256), (1, 1), (0, 0), (1, 1), ST)
run((128, 64, 256, 64), (64, 576), (3, 3), (1, 1), (1, 1), [("conv_halo", (0, 1))])
run((128, 64, 256, 64), None, (3, 3), (1, 1), (1, 1), [("wgrad_halo", (0, 1))], wgrad_gy=64)
run((128, 66, 258, 64), None, (3, 3), (0, 0), (2, 2), [("wgrad_halo", (1,))], wgrad_gy=128)
run((64, 66, 258, 128), None, (3, 3), (0, 0), (2, 2), [("wgrad_halo", (1,))], wgrad_gy=128)
run((128, 32, 128, 64), None, (1, 1), (0, 0), (1, 1), [("wgrad_halo", (1,))], wgrad_gy=128)
